"""Synthetic graphs of the reference's benchmark shapes, and the reference's on-disk format.

The reference trains on four whitespace-separated text files per dataset
(`features.txt row_ptr.txt col_idx.txt labels.txt`, /root/reference/README.md:20-27, loader
GATv2_edge_based.cu:24-64).  The Google-Drive datasets are not available offline, so every
config of BASELINE.json is generated here with fixed seeds.  The CSR is
destination-major (row = destination, col_idx = source, GATv2_edge_based.cu:67-84), rows are
sorted by source id, there are no duplicate edges, and every node has a self-loop so that every
softmax segment is non-empty (the reference is undefined for empty rows).
"""
import os

import numpy as np

# name -> N, E (CSR entries incl. one self-loop per node), feats, classes, model flags
CONFIGS = {
    "sample": dict(N=64, E=512, I=16, C=4, heads=[8, 1], outdims=[8, 8], optimizer="adam",
                   lr=0.01, clip=False, graph="uniform", feat="uniform", seed=1001),
    "cora": dict(N=2708, E=5429, I=1433, C=7, heads=[8, 1], outdims=[8, 8], optimizer="sgd",
                 lr=1e-4, clip=False, graph="uniform", feat="sparse_binary", density=0.013,
                 seed=1002),
    "pubmed": dict(N=19717, E=44338, I=500, C=3, heads=[8, 1], outdims=[8, 8], optimizer="sgd",
                   lr=1e-4, clip=True, graph="uniform", feat="sparse_pos", density=0.10,
                   seed=1003),
    "arxiv": dict(N=169343, E=1166243, I=128, C=40, heads=[4, 4, 1], outdims=[64, 64, 64],
                  optimizer="adam", lr=0.01, clip=False, graph="rmat", rmat=(0.57, 0.19, 0.19),
                  feat="normal", seed=1004),
    # R-MAT (0.45, 0.22, 0.22): max in-degree ~1e4 at this size, like ogbn-products' 17k
    "products": dict(N=2450000, E=61900000, I=100, C=47, heads=[4, 4, 1],
                     outdims=[128, 128, 128], optimizer="sgd", lr=1e-4, clip=False, graph="rmat",
                     rmat=(0.45, 0.22, 0.22), feat="normal", seed=1005),
}


def _rmat_pairs(rng, n_nodes, m, abc):
    """m (dst, src) pairs from an R-MAT recursion on 2^k >= n_nodes ids, folded mod n_nodes and
    relabelled by a random permutation so that hubs are not the low ids."""
    a, b, c = abc
    k = max(1, int(np.ceil(np.log2(n_nodes))))
    dst = np.zeros(m, np.int64)
    src = np.zeros(m, np.int64)
    for _ in range(k):
        r = rng.random(m, dtype=np.float32)
        # quadrants: a -> (0,0), b -> (0,1), c -> (1,0), d -> (1,1)   (dst bit, src bit)
        db = r >= (a + b)
        sb = ((r >= a) & (r < a + b)) | (r >= (a + b + c))
        dst = (dst << 1) | db
        src = (src << 1) | sb
    perm = rng.permutation(n_nodes).astype(np.int64)
    return perm[dst % n_nodes], perm[src % n_nodes]


def make_graph(N, E, kind="uniform", seed=0, rmat=(0.57, 0.19, 0.19)):
    """Destination-major CSR with exactly E entries: N self-loops + (E-N) distinct random edges."""
    assert E >= N and E - N <= N * (N - 1)
    rng = np.random.default_rng(seed)
    want = E - N
    keys = np.empty(0, np.int64)
    while len(keys) < want:
        m = int((want - len(keys)) * 1.15) + 64
        if kind == "rmat":
            d, s = _rmat_pairs(rng, N, m, rmat)
        else:
            d = rng.integers(0, N, m, dtype=np.int64)
            s = rng.integers(0, N, m, dtype=np.int64)
        ok = d != s
        keys = np.unique(np.concatenate([keys, d[ok] * N + s[ok]]))
    if len(keys) > want:
        keep = np.sort(rng.permutation(len(keys))[:want])
        keys = keys[keep]
    loops = np.arange(N, dtype=np.int64) * (N + 1)
    keys = np.sort(np.concatenate([keys, loops]))
    dst = keys // N
    col_idx = (keys - dst * N).astype(np.int32)
    row_ptr = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(dst, minlength=N), out=row_ptr[1:])
    return row_ptr.astype(np.int32), col_idx


def make_features(N, I, kind="normal", seed=0, density=0.1):
    rng = np.random.default_rng(seed + 7)
    if kind == "uniform":
        return rng.uniform(-1.0, 1.0, (N, I)).astype(np.float32)
    if kind == "normal":
        return rng.standard_normal((N, I), dtype=np.float32)
    mask = rng.random((N, I), dtype=np.float32) < density
    mask[np.arange(N), rng.integers(0, I, N)] = True  # no empty rows
    X = mask.astype(np.float32)
    if kind == "sparse_pos":
        X *= rng.random((N, I), dtype=np.float32) + 0.05
    X /= X.sum(axis=1, keepdims=True)
    return X.astype(np.float32)


def make_labels(N, C, seed=0):
    rng = np.random.default_rng(seed + 13)
    y = rng.integers(0, C, N).astype(np.int32)
    y[N - 1] = C - 1  # the reference derives C as max(label)+1 (GATv2_edge_based.cu:1106-1107)
    return y


def make_dataset(name, scale=1.0):
    """One of CONFIGS; scale < 1 shrinks N and E together (same mean degree, same feature and
    class counts) for bounded CPU / reference-binary samples of the big shapes."""
    cfg = dict(CONFIGS[name])
    N = max(8, int(round(cfg["N"] * scale)))
    E = max(N, int(round(cfg["E"] * scale)))
    E = min(E, N + N * (N - 1))
    row_ptr, col_idx = make_graph(N, E, cfg["graph"], cfg["seed"], cfg.get("rmat", (0.57, 0.19, 0.19)))
    X = make_features(N, cfg["I"], cfg["feat"], cfg["seed"], cfg.get("density", 0.1))
    y = make_labels(N, cfg["C"], cfg["seed"])
    cfg.update(N=N, E=E, name=name, scale=scale)
    return dict(row_ptr=row_ptr, col_idx=col_idx, X=X, labels=y, cfg=cfg)


def init_params(heads, outdims, I0, C, seed=0):
    """Xavier-uniform with the reference's limits: sqrt(6/(2*in+out)) for W and a
    (GATv2_edge_based.cu:208), sqrt(6/(C+out_last)) for W_o (GATv2_edge_based.cu:236)."""
    rng = np.random.default_rng(seed + 101)
    Ws, As = [], []
    I = I0
    for H, D in zip(heads, outdims):
        lim = np.sqrt(6.0 / (2 * I + D))
        Ws.append(rng.uniform(-lim, lim, (H * D, 2 * I)).astype(np.float32))
        As.append(rng.uniform(-lim, lim, (H * D,)).astype(np.float32))
        I = H * D
    lim = np.sqrt(6.0 / (C + outdims[-1]))
    Wo = rng.uniform(-lim, lim, (C, outdims[-1])).astype(np.float32)
    return Ws, As, Wo


# ------------------------------------------------------------------ reference text format
def write_txt(path, ds):
    """Writes the reference's four files.  Floats are printed with 9 significant digits so that
    the reference's `iss >> float` (GATv2_edge_based.cu:36) recovers the exact fp32 values."""
    os.makedirs(path, exist_ok=True)
    np.savetxt(os.path.join(path, "features.txt"), ds["X"], fmt="%.9g", delimiter=" ")
    for key, fn in (("row_ptr", "row_ptr.txt"), ("col_idx", "col_idx.txt"), ("labels", "labels.txt")):
        np.savetxt(os.path.join(path, fn), ds[key], fmt="%d")


def read_txt(path):
    """Same semantics as the reference loader: N = number of feature lines, I = tokens per line
    (must be constant), row_ptr must have N+1 entries and labels N (GATv2_edge_based.cu:24-64,
    1079-1100)."""
    X = np.loadtxt(os.path.join(path, "features.txt"), dtype=np.float32, ndmin=2)
    row_ptr = np.loadtxt(os.path.join(path, "row_ptr.txt"), dtype=np.int64).astype(np.int32).ravel()
    col_idx = np.loadtxt(os.path.join(path, "col_idx.txt"), dtype=np.int64).astype(np.int32).ravel()
    labels = np.loadtxt(os.path.join(path, "labels.txt"), dtype=np.int64).astype(np.int32).ravel()
    if len(row_ptr) != X.shape[0] + 1:
        raise ValueError("Invalid row_ptr length")
    if len(labels) != X.shape[0]:
        raise ValueError("Invalid labels length")
    return dict(row_ptr=row_ptr, col_idx=col_idx, X=X, labels=labels)
