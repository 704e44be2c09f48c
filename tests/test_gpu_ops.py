"""Op-level C-ABI entry points (include/gatx.h) on host buffers against the oracle's layer-level functions:
one kernel family at a time, without an epoch around it.

  gatx_op_edge_fwd   vs orc.layer_forward   (EB:279-459)   score / alpha / hpre / Hout, 2e-5
  gatx_op_edge_bwd   vs orc.layer_backward  (EB:612-798)   g_pre / gP_l / gP_r / ga / ge, 2e-4
  gatx_op_softmax_ce vs orc.head / loss     (EB:132-141, 514-550, 566-572)   y 2e-6, predicted labels bit-exact
  gatx_op_optimizer  vs orc_clip_grad_norm / orc_adam / orc_sgd (EB:250-278, 896-923), 2e-6
"""
import ctypes as C

import numpy as np
import pytest

import datasets
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gatx():
    import gatx as g
    g.load()
    return g


# (heads, outdim): warp-per-row, streaming 512-float rows, pair (1 x 128), generic scalar, pair with 64-float rows
FAMILIES = [(8, 8), (4, 128), (1, 128), (3, 5), (1, 64)]


@pytest.mark.parametrize("H,D", FAMILIES, ids=["narrow", "stream", "pair", "generic", "pair64"])
def test_op_edge_forward_and_backward(gatx, orc, H, D):
    N, E = 1500, 14000
    rp, ci = datasets.make_graph(N, E, "rmat", 31)
    rng = np.random.default_rng(H * 100 + D)
    F = H * D
    Pl = rng.standard_normal((N, F), dtype=np.float32)
    Pr = rng.standard_normal((N, F), dtype=np.float32)
    a = (rng.standard_normal(F) * 0.3).astype(np.float32)
    ref = orc.layer_forward(rp, ci, H, D, Pl, Pr, a, False)
    out = gatx.op_edge_fwd(rp, ci, H, D, Pl, Pr, a)
    assert rel_err(out["score"].T, ref["score"]) < 2e-5
    assert np.abs(out["alpha"].T - ref["alpha"]).max() < 1e-5
    assert rel_err(out["hpre"], ref["hpre"]) < 4e-5
    assert rel_err(out["Hout"], ref["Hout"]) < 4e-5
    gHout = rng.standard_normal((N, F), dtype=np.float32)
    g_pre = (gHout * np.where(ref["hpre"] > 0, 1.0, 0.01)).astype(np.float32)  # EB:879-893
    rb = orc.layer_backward(rp, ci, H, D, np.zeros((N, 1), np.float32), np.zeros((F, 2), np.float32), a, Pl, Pr,
                            ref["alpha"], g_pre, want_gx=False)
    bw = gatx.op_edge_bwd(rp, ci, H, D, Pl, Pr, a, gHout)
    assert rel_err(bw["g_pre"], g_pre) < 2e-6
    assert rel_err(bw["ge"].T, rb["ge"]) < 2e-4
    assert rel_err(bw["gPr"], rb["gPr"]) < 2e-4
    assert rel_err(bw["gPl"], rb["gPl"]) < 2e-4
    assert rel_err(bw["ga"], rb["ga"], floor=1e-2) < 2e-4


def test_op_softmax_ce(gatx, orc):
    rng = np.random.default_rng(3)
    N, Cc = 5000, 47
    z = (rng.standard_normal((N, Cc)) * 3).astype(np.float32)
    z[17, 5] = z[17, 9] = z[17].max() + 1.0  # a tie: first maximum wins (EB:530-535)
    labels = rng.integers(0, Cc, N).astype(np.int32)
    mask = (rng.random(N) < 0.6).astype(np.uint8)
    zz = z.astype(np.float64) - z.max(1, keepdims=True)
    y_ref = np.exp(zz) / (np.exp(zz).sum(1, keepdims=True) + 1e-8)
    la = orc.loss_acc(y_ref.astype(np.float32), labels)
    for m in (None, mask):
        out = gatx.op_softmax_ce(z, labels, m)
        assert np.abs(out["y"] - y_ref).max() < 2e-6
        assert np.array_equal(out["pred"], la["pred"]) and out["pred"][17] == 5
        sel = np.ones(N, bool) if m is None else m.astype(bool)
        assert abs(out["loss_sum"] - la["losses"][sel].astype(np.float64).sum()) < 1e-4 * sel.sum()
        assert out["correct"] == int(la["correct"][sel].sum())
        onehot = np.zeros((N, Cc), np.float32)
        onehot[np.arange(N), labels] = 1.0
        dz_ref = (y_ref - onehot) * sel[:, None]  # EB:572: sum-loss gradient; masked-out nodes contribute nothing
        assert np.abs(out["dz"] - dz_ref).max() < 2e-6
    with pytest.raises(gatx.GatxError):
        gatx.op_softmax_ce(z, np.full(N, Cc, np.int32))


@pytest.mark.parametrize("optimizer,clip", [("adam", False), ("adam", True), ("sgd", True)])
def test_op_optimizer(gatx, orc, optimizer, clip):
    rng = np.random.default_rng(11)
    n, ends = 20000, (15000, 15800, 20000)
    p0 = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pr, mr, vr = p0.copy(), m.copy(), v.copy()
    p = p0.copy()
    lib = orc.lib()
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    for t in (1, 2, 3):
        g = (rng.standard_normal(n) * (0.5 if t == 2 else 0.01)).astype(np.float32)  # epoch 2 trips the 5.0 threshold
        gr = g.copy()
        if clip:
            b = 0
            for e in ends:
                seg = gr[b:e]  # a view: clipped in place
                lib.orc_clip_grad_norm(fp(seg), C.c_int64(e - b), C.c_float(5.0))
                b = e
        if optimizer == "adam":
            lib.orc_adam(fp(pr), fp(gr), fp(mr), fp(vr), C.c_float(1e-2), C.c_int64(n), C.c_float(0.9), C.c_float(0.999),
                         C.c_float(1e-8), t)
        else:
            lib.orc_sgd(fp(pr), fp(gr), C.c_float(1e-2), C.c_int64(n))
        p, gz, m2, v2 = gatx.op_optimizer(p, g, ends, optimizer, clip, lr=1e-2, t=t, m=m, v=v)
        if optimizer == "adam":
            m, v = m2, v2
            assert rel_err(m, mr) < 2e-6 and rel_err(v, vr) < 2e-6
        assert not gz.any()  # gradients are reset by the update (EB:1631-1633)
        assert rel_err(p, pr) < 2e-6, t
