"""Small end-to-end run for compute-sanitizer (GPU box): narrow-row kernels, streaming kernels with tiny chunks and a hub,
both GEMM modes.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("graph-attention-network-gatv2-_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
os.environ["GATX_CHUNK"] = "64"
import gatx
from helpers import make_engine, make_problem
for shape in [(120, 700, 9, 3, (8, 1), (8, 8), None), (400, 3000, 12, 4, (4, 2, 1), (32, 128, 128), 300)]:
    N, E, I, C, heads, outdims, hub = shape
    p = make_problem(N, E, I, C, heads, outdims, "rmat", seed=N, hub=hub)
    for mode in (0, 1):
        eng = make_engine(gatx, p, gemm_mode=mode, optimizer="adam", lr=0.01, clip=True, keep_debug=(mode == 1))
        for t in (1, 2):
            print(shape[:2], mode, eng.train_epoch(t))
        eng.close()
print("done")
