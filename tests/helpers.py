import numpy as np

import datasets


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max(max |b|, floor); `floor` keeps an exactly-zero reference comparable."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def make_problem(N, E, I, C, heads, outdims, kind="uniform", seed=0, wscale=2.0, hub=None):
    row_ptr, col_idx = datasets.make_graph(N, E, kind, seed)
    if hub:  # add a destination and a source with `hub` edges (exercises the CTA-per-row kernels)
        rng = np.random.default_rng(seed + 5)
        dst = np.repeat(np.arange(N), np.diff(row_ptr))
        src = col_idx.astype(np.int64)
        extra_s = rng.choice(N, hub, replace=False)
        extra_d = rng.choice(N, hub, replace=False)
        d = np.concatenate([dst, np.full(hub, 3), extra_d])
        s = np.concatenate([src, extra_s, np.full(hub, 5)])
        keys = np.unique(d.astype(np.int64) * N + s)
        d = keys // N
        col_idx = (keys - d * N).astype(np.int32)
        row_ptr = np.zeros(N + 1, np.int64)
        np.cumsum(np.bincount(d, minlength=N), out=row_ptr[1:])
        row_ptr = row_ptr.astype(np.int32)
    X = datasets.make_features(N, I, "uniform", seed)
    y = datasets.make_labels(N, C, seed)
    Ws, As, Wo = datasets.init_params(heads, outdims, I, C, seed)
    Ws = [w * wscale for w in Ws]
    As = [a * wscale for a in As]
    return dict(row_ptr=row_ptr, col_idx=col_idx, X=X, labels=y, Ws=Ws, As=As, Wo=Wo, heads=list(heads),
                outdims=list(outdims), C=C)


def make_engine(gatx, p, **kw):
    eng = gatx.Engine(p["heads"], p["outdims"], **kw)
    eng.set_graph(p["row_ptr"], p["col_idx"])
    eng.set_features(p["X"])
    eng.set_labels(p["labels"], p["C"])
    for l in range(len(p["heads"])):
        eng.set_params(l, p["Ws"][l], p["As"][l])
    eng.set_wo(p["Wo"])
    return eng


def make_oracle(orc, p, **kw):
    m = orc.Model(p["heads"], p["outdims"], p["row_ptr"], p["col_idx"], p["X"], p["labels"], num_classes=p["C"], **kw)
    for l in range(len(p["heads"])):
        m.set_params(l, p["Ws"][l], p["As"][l])
    m.set_wo(p["Wo"])
    return m
