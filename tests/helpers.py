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


def sub_problem(row_ptr, col_idx, rows):
    """Compact sub-graph holding the COMPLETE in-edge segments of the destination rows `rows` (global ids).

    Used by the full-size parity tests: the oracle cannot run a 61.9 M-edge layer in seconds, but it can run the
    layer on the sub-graph induced by a sample of destination rows when it is fed the engine's own projected
    features.  Returns dict(nodes, ptr, col, eidx, rows_local):
      nodes      sorted global ids of the sub-graph's nodes (the sampled rows and every source they gather)
      ptr, col   destination-major CSR over len(nodes) compact ids; rows that were not sampled are empty
      eidx       global CSR position of every sub-graph edge, in sub-graph edge order
      rows_local compact ids of the sampled rows
    A source's gP_l from the sub-graph is complete iff all of its out-edges end in sampled rows."""
    row_ptr = np.asarray(row_ptr, np.int64)
    rows = np.unique(np.asarray(rows, np.int64))
    deg = row_ptr[rows + 1] - row_ptr[rows]
    starts = np.repeat(row_ptr[rows] - np.concatenate([[0], np.cumsum(deg)[:-1]]), deg)
    eidx = starts + np.arange(int(deg.sum()), dtype=np.int64)
    srcs = np.asarray(col_idx)[eidx].astype(np.int64)
    nodes = np.unique(np.concatenate([rows, srcs]))
    rows_local = np.searchsorted(nodes, rows)
    sub_deg = np.zeros(len(nodes), np.int64)
    sub_deg[rows_local] = deg
    ptr = np.zeros(len(nodes) + 1, np.int64)
    np.cumsum(sub_deg, out=ptr[1:])
    return dict(nodes=nodes, ptr=ptr.astype(np.int32), col=np.searchsorted(nodes, srcs).astype(np.int32), eidx=eidx,
                rows_local=rows_local)


def out_edge_closure(row_ptr, col_idx, sources):
    """Destination rows reached by the out-edges of `sources` (global ids): sampling these rows makes the sub-graph's
    gP_l complete for every node of `sources`."""
    N = len(row_ptr) - 1
    dst = np.repeat(np.arange(N, dtype=np.int64), np.diff(np.asarray(row_ptr, np.int64)))
    return np.unique(dst[np.isin(np.asarray(col_idx), np.asarray(sources))])
