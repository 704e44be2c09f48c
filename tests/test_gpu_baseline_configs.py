"""GPU parity at the LITERAL small configs of BASELINE.json (configs 2 and 3; shapes from the reference's
README.md:30-34, flags from BASELINE.md section 3):

  cora-shaped    2 708 nodes,  5 429 edges, 1 433 features, 7 classes, --heads 8,1 --outdims 8,8, SGD
  pubmed-shaped 19 717 nodes, 44 338 edges,   500 features, 3 classes, --heads 8,1 --outdims 8,8, SGD --clip

Each runs forward (EB:279-550), backward (EB:553-893) and a 10-epoch loss curve (EB:1370-1642, clip EB:1561-1566)
against the CPU oracle in BOTH GEMM modes: fp32 CUDA cores (2e-5 forward / 2e-4 gradients, predicted labels
bit-exact) and TF32 tensor cores (5e-3 / 5e-2; K = 1 433 is padded to 1 436 for the 16-byte row pitch of the TMA
maps).  Tolerances are max |a-b| / max |b| as in test_gpu_parity.py.
"""
import numpy as np
import pytest

from helpers import make_engine, make_oracle, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = {1: 2e-5, 0: 5e-3}
BWD_TOL = {1: 2e-4, 0: 5e-2}


@pytest.fixture(scope="module")
def gatx():
    import gatx as g
    g.load()
    return g


def _problem(name, seed=11):
    import datasets
    ds = datasets.make_dataset(name)
    cfg = ds["cfg"]
    assert (cfg["N"], cfg["E"]) == {"cora": (2708, 5429), "pubmed": (19717, 44338)}[name]
    Ws, As, Wo = datasets.init_params(cfg["heads"], cfg["outdims"], cfg["I"], cfg["C"], seed)
    return cfg, dict(row_ptr=ds["row_ptr"], col_idx=ds["col_idx"], X=ds["X"], labels=ds["labels"], Ws=Ws, As=As, Wo=Wo,
                     heads=cfg["heads"], outdims=cfg["outdims"], C=cfg["C"])


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("name", ["cora", "pubmed"])
def test_forward_backward_at_baseline_config(gatx, orc, name, mode):
    cfg, p = _problem(name)
    eng = make_engine(gatx, p, gemm_mode=mode, keep_debug=True, optimizer=cfg["optimizer"], lr=cfg["lr"],
                      clip=cfg["clip"])
    ref = make_oracle(orc, p, optimizer=cfg["optimizer"], lr=cfg["lr"], clip=cfg["clip"])
    eng.forward()
    loss, acc = eng.loss_acc()
    ref.forward()
    rl = ref.loss()
    ft, bt = FWD_TOL[mode], BWD_TOL[mode]
    E = len(p["col_idx"])
    for l in range(2):
        H = cfg["heads"][l]
        assert rel_err(eng.tensor(gatx.T_PL, l), ref.tensor(orc.T_PL, l).ravel()) < ft, ("Pl", l)
        assert rel_err(eng.tensor(gatx.T_PR, l), ref.tensor(orc.T_PR, l).ravel()) < ft, ("Pr", l)
        assert rel_err(eng.tensor(gatx.T_SCORE, l).reshape(E, H).T, ref.tensor(orc.T_SCORE, l)) < ft, ("score", l)
        assert np.abs(eng.tensor(gatx.T_ALPHA, l).reshape(E, H).T - ref.tensor(orc.T_ALPHA, l)).max() < 5 * ft, ("alpha", l)
        assert rel_err(eng.tensor(gatx.T_HOUT, l), ref.tensor(orc.T_HOUT, l).ravel()) < 2 * ft, ("Hout", l)
    assert np.abs(eng.tensor(gatx.T_Y) - ref.tensor(orc.T_Y).ravel()).max() < 5 * ft
    assert abs(loss - rl["avg"]) < max(5 * ft, 1e-5) * max(1.0, abs(rl["avg"]))
    if mode == 1:
        assert np.array_equal(eng.tensor(gatx.T_PRED), rl["pred"])  # bit-exact predicted labels (EB:530-535)
        assert acc == pytest.approx(rl["acc"], abs=1e-7)
    else:
        assert (eng.tensor(gatx.T_PRED) != rl["pred"]).mean() < 0.02  # near-tie logits only
    eng.backward()
    ref.backward()
    for l in range(2):
        gh, gh_ref = eng.tensor(gatx.T_GH, l), ref.tensor(orc.T_GH, l).ravel()
        if mode == 1:
            assert rel_err(gh, gh_ref) < bt, ("g_h", l)
        else:
            assert np.linalg.norm(gh - gh_ref) < bt * np.linalg.norm(gh_ref), ("g_h L2", l)
        gW, gW_ref = eng.tensor(gatx.T_GW, l), ref.tensor(orc.T_GW, l).ravel()
        if mode == 1:
            assert rel_err(gW, gW_ref) < bt, ("gW", l)
        else:
            # TF32 with SPARSE features: a column of X has ~35 non-zeros (cora), so one gW element is a sum of ~35 terms
            # and a single LeakyReLU' flip (a pre-activation that TF32 rounding moves across 0 changes that gradient
            # element by 99 %) shows up undamped in the maximum norm -- measured 9e-2 of the maximum on the cora shape
            # against 1e-3 on a typical element.  The L2 norm states the TF32 error, the maximum norm bounds outliers.
            assert np.linalg.norm(gW - gW_ref) < bt * np.linalg.norm(gW_ref), ("gW L2", l)
            assert rel_err(gW, gW_ref) < 0.2, ("gW max", l)
        assert rel_err(eng.tensor(gatx.T_GA, l), ref.tensor(orc.T_GA, l).ravel(), floor=1e-2) < bt, ("ga", l)
    assert rel_err(eng.tensor(gatx.T_GWO), ref.tensor(orc.T_GWO).ravel()) < bt
    # the update itself: clip (pubmed) + SGD, then the parameters
    eng.step(1)
    ref.step(1)
    wt = 2e-5 if mode == 1 else 5e-3  # W moves by lr * g, small against W itself: this checks the update rule
    for l in range(2):
        assert rel_err(eng.tensor(gatx.T_W, l), ref.tensor(orc.T_W, l).ravel()) < wt, ("W after step", l)
    assert rel_err(eng.tensor(gatx.T_WO), ref.tensor(orc.T_WO).ravel()) < wt
    eng.close()


@pytest.mark.parametrize("mode", [1, 0], ids=["fp32_simt", "tf32_tc"])
@pytest.mark.parametrize("name", ["cora", "pubmed"])
def test_loss_curve_at_baseline_config(gatx, orc, name, mode):
    """10 epochs with the config's own flags (the reference's default SGD step of 1e-4 on the SUMMED loss, EB:572;
    pubmed with --clip).  A second run at a 30x larger step makes the curve move visibly over the 10 epochs."""
    cfg, p = _problem(name)
    for lr in (cfg["lr"], 30 * cfg["lr"]):
        eng = make_engine(gatx, p, gemm_mode=mode, optimizer=cfg["optimizer"], lr=lr, clip=cfg["clip"])
        ref = make_oracle(orc, p, optimizer=cfg["optimizer"], lr=lr, clip=cfg["clip"])
        tol = 2e-4 if mode == 1 else 5e-3
        first = last = None
        for t in range(1, 11):
            gl, ga = eng.train_epoch(t)
            rl, ra = ref.epoch(t)
            assert abs(gl - rl) < tol * max(1.0, rl), (lr, t, gl, rl)
            # accuracy: a handful of near-tie nodes may flip in TF32 mode
            assert abs(ga - ra) <= (2.0 if mode == 1 else 0.02 * cfg["N"]) / cfg["N"] + 1e-6, (lr, t, ga, ra)
            first = rl if first is None else first
            last = rl
        assert last < first  # the curve is a training curve, not a constant
        for l in range(2):
            assert rel_err(eng.tensor(gatx.T_W, l), ref.tensor(orc.T_W, l).ravel()) < 50 * tol, ("W", l)
        eng.close()
