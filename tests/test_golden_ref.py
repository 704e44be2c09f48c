"""Pins the CPU oracle against the reference ITSELF: tests/golden/ref_edge_*.npz hold buffers dumped
by the reference's edge-based program (built from /root/reference by oracle/Makefile, run on a B200
by oracle/gen_golden.py with injected weights).  Tolerances absorb what the reference does
differently by construction: fp32 accumulation, `__expf`, and the arbitrary order of its float
atomics (EB:422, 579, 786, 793, 868-869)."""
import glob
import os

import numpy as np
import pytest

from helpers import rel_err

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_edge_*.npz")))


def load(path, orc):
    g = dict(np.load(path, allow_pickle=False))
    if "X" not in g:  # sparse feature fixtures (literal cora shape) store the non-zeros only
        X = np.zeros(int(np.prod(g["X_shape"])), np.float32)
        X[g["X_nnz_idx"]] = g["X_nnz_val"]
        g["X"] = X.reshape(tuple(g["X_shape"]))
    heads, outdims = g["heads"].tolist(), g["outdims"].tolist()
    m = orc.Model(heads, outdims, g["row_ptr"], g["col_idx"], g["X"], g["labels"], optimizer=str(g["optimizer"]),
                  clip=bool(g["clip"]), lr=float(g["lr"]))
    for l in range(len(heads)):
        m.set_params(l, g["W_%d" % l], g["a_%d" % l])
    m.set_wo(g["Wo"])
    return g, m, heads, outdims


def test_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_oracle_matches_reference_epoch1(path, orc):
    g, m, heads, outdims = load(path, orc)
    N, E = len(g["labels"]), len(g["col_idx"])
    # integer work: bit-exact
    src, dst = orc.csr_to_coo(g["row_ptr"], g["col_idx"])
    assert np.array_equal(src, g["ref_coo_src"]) and np.array_equal(dst, g["ref_coo_dst"])
    assert orc.max_degree(g["row_ptr"]) == int(g["max_degree"])
    assert orc.num_classes(g["labels"]) == int(g["num_classes"]) == m.C
    m.forward()
    loss = m.loss()
    m.backward()
    L = len(heads)
    # Reference defect D14 (found with the cora-shaped fixture): compute_output_gradients returns its out-of-range
    # threads (EB:566) BEFORE the block-wide copy of W_o into shared memory (EB:585-588), so in the last block only
    # the N % 128 surviving threads copy, and elements i >= N % 128 of W_o (i < C * D_L) stay uninitialised there.
    # When 0 < N % 128 < C * D_L the gradients g_h of the last N % 128 nodes are garbage, and with them everything
    # that sums over those nodes (ga, gW, and g_h of every node with an edge INTO them one layer down).  All of the
    # reference's real datasets are hit (cora 20 < 56, pubmed 5 < 24, arxiv 127 < 2560, products 16 < 6016).
    tail_nodes = N % 128
    d14 = 0 < tail_nodes < m.C * outdims[-1]
    bad = np.zeros(N, bool)
    if d14:
        bad[N - tail_nodes:] = True
    for l in range(L):
        assert rel_err(m.tensor(orc.T_SCORE, l).ravel(), g["ref_score_%d" % l]) < 2e-5, ("score", l)
        assert np.abs(m.tensor(orc.T_ALPHA, l).ravel() - g["ref_alpha_%d" % l]).max() < 2e-6, ("alpha", l)
        assert rel_err(m.tensor(orc.T_HPRE, l).ravel(), g["ref_hpre_%d" % l]) < 2e-5, ("hpre", l)
        assert rel_err(m.tensor(orc.T_HOUT, l).ravel(), g["ref_hout_%d" % l]) < 2e-5, ("hout", l)
    for l in range(L - 1, -1, -1):
        mine, ref = m.tensor(orc.T_GH, l), g["ref_gh_%d" % l].reshape(N, -1)
        assert rel_err(mine[~bad], ref[~bad]) < 1e-4, ("g_h", l)
        if d14:
            assert rel_err(mine[bad], ref[bad]) > 1e-2, "fixture no longer shows defect D14"
            # one layer down, the garbage reaches every source of an edge into a bad node (gP_l) besides the node itself
            rp, ci = g["row_ptr"], g["col_idx"]
            srcs = np.concatenate([ci[rp[i]:rp[i + 1]] for i in np.flatnonzero(bad)])
            bad = bad.copy()
            bad[srcs] = True
    assert np.abs(m.tensor(orc.T_Y).ravel() - g["ref_y"]).max() < 2e-6
    la = orc.loss_acc(m.tensor(orc.T_Y), m.labels)
    assert np.abs(la["losses"] - g["ref_loss"]).max() < 2e-5
    assert np.array_equal(la["correct"], g["ref_correct"])  # predicted-label hits, bit-exact
    assert abs(loss["avg"] - g["loss_curve"][0, 0]) < 2e-6 + 1e-6  # printed with %f
    assert rel_err(m.tensor(orc.T_GWO).ravel(), g["ref_gWo"]) < 1e-4  # computed before the defective copy (EB:576-581)
    if d14:
        assert "defect" in os.path.basename(path)
        return  # ga and gW sum over the corrupted nodes
    ga = np.concatenate([m.tensor(orc.T_GA, l).ravel() for l in range(L)])
    assert rel_err(ga, g["ref_ga"]) < 1e-4
    # gW layer by layer.  Reference defect D13: when 0 < E % 256 < 2*in_dim the threads of the last CTA
    # that returned early (EB:722) never zero their columns of sh_grad_w (EB:747-749); those columns of
    # the reference's gW are garbage and only the columns below E % 256 can be compared.
    off, tail, defect = 0, E % 256, False
    for l in range(L):
        mine = m.tensor(orc.T_GW, l)
        ref = g["ref_gW"][off:off + mine.size].reshape(mine.shape)
        off += mine.size
        if 0 < tail < mine.shape[1]:
            defect = True
            assert rel_err(mine[:, :tail], ref[:, :tail]) < 1e-4, ("gW", l)
            if "tail_block" in os.path.basename(path):  # (uninitialised shared memory: it may also happen to be zero)
                assert rel_err(mine[:, tail:], ref[:, tail:]) > 1e-3, "fixture no longer shows defect D13"
        else:
            assert rel_err(mine, ref) < 1e-4, ("gW", l)
    assert defect == ("defect" in os.path.basename(path))
    if defect:
        return
    # clip + optimizer (EB:1560-1625)
    m.step(1)
    W = np.concatenate([m.tensor(orc.T_W, l).ravel() for l in range(L)])
    a = np.concatenate([m.tensor(orc.T_A, l).ravel() for l in range(L)])
    lr, adam = float(g["lr"]), str(g["optimizer"]) == "adam"
    for mine, ref in ((W, g["ref_W_after"]), (a, g["ref_a_after"]), (m.tensor(orc.T_WO).ravel(), g["ref_Wo_after"])):
        if adam:
            # the first Adam step is lr * g / (|g| + 1e-8): for |g| ~ 1e-8 it amplifies the reference's
            # atomic-order noise up to +-lr, so a few elements may differ by up to 2*lr
            d = np.abs(mine - ref)
            assert d.max() <= 2 * lr * 1.001 and np.mean(d > 2e-5 * np.abs(ref).max()) < 0.01
        else:
            assert rel_err(mine, ref) < 2e-5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[9:-4] for p in GOLDEN])
def test_oracle_matches_reference_loss_curve(path, orc):
    """Whole loss/accuracy curve of the reference (with its never-zeroed aggregation buffer fixed,
    SURVEY D1).  The unpatched curve is kept in the fixture to document that defect."""
    if "defect" in os.path.basename(path):
        pytest.skip("gradients of this fixture are corrupted by reference defect D13")
    g, m, heads, outdims = load(path, orc)
    curve = g["loss_curve"]
    N = len(g["labels"])
    for t in range(1, len(curve) + 1):
        avg, acc = m.epoch(t)
        assert abs(avg - curve[t - 1, 0]) < 5e-4 * max(1.0, curve[t - 1, 0]), (t, avg, curve[t - 1])
        assert abs(100 * acc - curve[t - 1, 1]) <= 100.0 / N + 0.006, (t, acc, curve[t - 1])
    d1 = g["loss_curve_unpatched_d1"]
    assert abs(d1[0, 0] - curve[0, 0]) < 1e-6 and abs(d1[2, 0] - curve[2, 0]) > 1e-2
