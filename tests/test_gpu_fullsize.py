"""Parity at BASELINE.json's full sizes.

* arxiv shape (169 343 nodes, 1 166 243 edges, 3 layers 4,4,1 x 64): the whole forward + backward against the
  CPU oracle (fp32 CUDA-core GEMM mode, tolerances of test_gpu_parity.py).
* products shape (2.45 M nodes, 61.9 M edges, 3 layers 4,4,1 x 128): size-independent properties of the fused
  edge kernels, plus the oracle on a random sample of destination rows fed with the engine's own projections:
    - every softmax segment sums to 1;
    - sampled rows: alpha and the layer output equal the oracle's;
    - softmax backward: ge sums to 0 over every segment;
    - checksum of checksums: colsum(gP_l) - colsum(gP_r) = colsum(g_h)   (because sum_seg alpha = 1);
    - two runs give bit-identical losses (no atomics).
"""
import os
import sys

import numpy as np
import pytest

from helpers import make_engine, make_oracle, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gatx():
    import gatx as g
    g.load()
    return g


def _dataset(name):
    import datasets
    cache = os.path.join("/dev/shm", "gatx_%s_1" % name)
    if os.path.exists(os.path.join(cache, "done")):  # bench.py's cache
        ds = {k: np.load(os.path.join(cache, k + ".npy")) for k in ("row_ptr", "col_idx", "X", "labels")}
        ds["cfg"] = dict(datasets.CONFIGS[name])
        return ds
    return datasets.make_dataset(name)


def test_arxiv_full_epoch_vs_oracle(gatx, orc):
    import datasets
    ds = _dataset("arxiv")
    cfg = ds["cfg"]
    Ws, As, Wo = datasets.init_params(cfg["heads"], cfg["outdims"], cfg["I"], cfg["C"], 5)
    p = dict(row_ptr=ds["row_ptr"], col_idx=ds["col_idx"], X=ds["X"], labels=ds["labels"], Ws=Ws, As=As, Wo=Wo,
             heads=cfg["heads"], outdims=cfg["outdims"], C=cfg["C"])
    eng = make_engine(gatx, p, gemm_mode=gatx.GEMM_FP32_SIMT, optimizer="adam", lr=0.01)
    ref = make_oracle(orc, p, optimizer="adam", lr=0.01)
    eng.forward()
    loss, acc = eng.loss_acc()
    ref.forward()
    rl = ref.loss()
    assert rel_err(eng.tensor(gatx.T_HOUT, 2), ref.tensor(orc.T_HOUT, 2).ravel()) < 1e-4
    assert abs(loss - rl["avg"]) < 1e-5 * max(1.0, rl["avg"])
    assert (eng.tensor(gatx.T_PRED) != rl["pred"]).sum() <= 2  # fp32 argmax near-ties on 169 343 x 40 logits
    eng.backward()
    ref.backward()
    # fp32 sums over 1.17 M edges / 169 k nodes against the oracle's fp64 accumulation: 2e-3 of the maximum
    for l in range(3):
        assert rel_err(eng.tensor(gatx.T_GW, l), ref.tensor(orc.T_GW, l).ravel()) < 2e-3, l
        assert rel_err(eng.tensor(gatx.T_GA, l), ref.tensor(orc.T_GA, l).ravel(), floor=1e-2) < 2e-3, l
    assert rel_err(eng.tensor(gatx.T_GWO), ref.tensor(orc.T_GWO).ravel()) < 2e-3
    eng.close()


def test_products_full_size_properties(gatx, orc):
    import datasets
    ds = _dataset("products")
    cfg = ds["cfg"]
    N, E = len(ds["labels"]), len(ds["col_idx"])
    assert (N, E) == (2450000, 61900000)
    rp, ci = np.asarray(ds["row_ptr"]), np.asarray(ds["col_idx"])
    losses = []
    for run in range(2):
        eng = gatx.Engine(cfg["heads"], cfg["outdims"], optimizer="adam", lr=0.001, keep_debug=(run == 0))
        eng.set_graph(rp, ci)
        eng.set_features(np.asarray(ds["X"]))
        eng.set_labels(np.asarray(ds["labels"]), cfg["C"])
        eng.init_params(99)
        eng.forward()
        losses.append(eng.loss_acc())
        if run == 1:
            eng.close()
            break
        # --- forward properties, layer 0 (4 heads x 128) ---
        H, D = cfg["heads"][0], cfg["outdims"][0]
        alpha = eng.tensor(gatx.T_ALPHA, 0).reshape(E, H)
        seg = np.add.reduceat(alpha.astype(np.float64), rp[:-1].astype(np.int64), axis=0)
        assert np.abs(seg - 1.0).max() < 2e-5  # every destination has a self-loop: no empty segment
        rng = np.random.default_rng(0)
        hub = int(np.argmax(np.diff(rp)))
        rows = np.unique(np.concatenate([rng.choice(N, 1500, replace=False), [hub, 0, N - 1]]))
        deg = np.diff(rp)[rows]
        sub_ptr = np.zeros(len(rows) + 1, np.int32)
        np.cumsum(deg, out=sub_ptr[1:])
        eidx = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in rows])
        Pl = eng.tensor(gatx.T_PL, 0).reshape(N, H * D)
        Pr = eng.tensor(gatx.T_PR, 0).reshape(N, H * D)[rows]
        a0 = eng.tensor(gatx.T_A, 0)
        out = orc.layer_forward(sub_ptr, np.ascontiguousarray(ci[eidx]), H, D, Pl, np.ascontiguousarray(Pr), a0, False)
        assert np.abs(alpha[eidx].T - out["alpha"]).max() < 1e-5
        Hout = eng.tensor(gatx.T_HOUT, 0).reshape(N, H * D)[rows]
        assert rel_err(Hout, out["Hout"]) < 2e-5
        del Pl, alpha
        # --- backward properties ---
        eng.backward()
        ge = eng.tensor(gatx.T_GE, 0).reshape(E, H).astype(np.float64)
        seg0 = np.add.reduceat(ge, rp[:-1].astype(np.int64), axis=0)
        segabs = np.add.reduceat(np.abs(ge), rp[:-1].astype(np.int64), axis=0)
        assert np.abs(seg0).max() < 1e-4 * max(segabs.max(), 1e-30)
        del ge
        for l in (2, 0):
            Hl = cfg["heads"][l]
            F = Hl * cfg["outdims"][l]
            gh = eng.tensor(gatx.T_GH, l).reshape(N, F).astype(np.float64).sum(0)
            gpl = eng.tensor(gatx.T_GPL, l).reshape(N, F).astype(np.float64).sum(0)
            gpr = eng.tensor(gatx.T_GPR, l).reshape(N, F).astype(np.float64).sum(0)
            scale = np.abs(gh).max() + np.abs(gpl).max()
            assert np.abs(gpl - gpr - gh).max() < 2e-4 * scale, l
        eng.close()
    assert losses[0] == losses[1]
    assert abs(losses[0][0] - np.log(cfg["C"])) < 0.5  # Xavier init: loss near log(47)
