"""Debug helper (GPU box): decode how the MN-major tcgen05 kernel interprets its operands."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-attention-network-gatv2-_b200"))
import numpy as np
import gatx
np.set_printoptions(linewidth=250, threshold=100000)
M, N, K = 128, 64, 32
k = np.arange(K)[:, None]
# exp1: A[k][m] = m, B[k][n] = [n == k]  ->  C[m][n] = m for n < K
A = np.tile(np.arange(M, dtype=np.float32), (K, 1)); B = (np.arange(N)[None, :] == k).astype(np.float32)
C = gatx.op_gemm(A, B, form=1, mode=0); ref = A.T @ B
print("exp1 (A=m, B=delta) max err", np.abs(C - ref).max())
if np.abs(C - ref).max() > 0:
    print("C[:,0:4].T", C[:, 0:4].T.astype(int)); print("C[0:4,:]", C[0:4, :].astype(int))
# exp2: A[k][m] = k, B = delta -> C[m][n] = n (n<K)
A = np.tile(np.arange(K, dtype=np.float32)[:, None], (1, M))
C = gatx.op_gemm(A, B, form=1, mode=0); ref = A.T @ B
print("exp2 (A=k, B=delta) max err", np.abs(C - ref).max())
if np.abs(C - ref).max() > 0:
    print("C[0:4,:]", C[0:4, :].astype(int)); print("C[:,0:4].T", C[:, 0:4].T.astype(int))
# exp3: B[k][n] = n, A[k][m] = [m == k] -> C[m][n] = n for m < K
A = (np.arange(M)[None, :] == k).astype(np.float32); B = np.tile(np.arange(N, dtype=np.float32), (K, 1))
C = gatx.op_gemm(A, B, form=1, mode=0); ref = A.T @ B
print("exp3 (A=delta, B=n) max err", np.abs(C - ref).max())
if np.abs(C - ref).max() > 0:
    print("C[0:4,:]", C[0:4, :].astype(int)); print("C[:,0:4].T", C[:, 0:4].T.astype(int))
rng = np.random.default_rng(0)
for (M, N, K) in [(128, 64, 32), (128, 64, 64), (128, 256, 320), (64, 16, 64), (512, 128, 3000), (16, 8, 100000)]:
    A = rng.standard_normal((K, M)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    C = gatx.op_gemm(A, B, form=1, mode=0)
    print((M, N, K), "rel err", np.abs(C - ref).max() / np.abs(ref).max())
