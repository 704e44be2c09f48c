"""Diagnostics for the 64-float pair kernels: every tensor of a small multi-chunk case and the arxiv last layer against the
oracle, with the worst rows / edges located.  python tools/debug_pair64.py small|arxiv"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("graph-attention-network-gatv2-_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gatx
import orc
from helpers import make_engine, make_oracle, make_problem, rel_err

gatx.load()
which = sys.argv[1]
if which == "small":
    p = make_problem(900, 7000, 10, 4, (4, 2, 1), (32, 128, int(sys.argv[2]) if len(sys.argv) > 2 else 64), "rmat", seed=21, hub=600)
else:  # "arxiv" / "race"
    import datasets
    ds = datasets.make_dataset("arxiv")
    cfg = ds["cfg"]
    Ws, As, Wo = datasets.init_params(cfg["heads"], cfg["outdims"], cfg["I"], cfg["C"], 5)
    if len(sys.argv) > 3:  # race <runs> <last outdim>: the same graph with a 128-float last layer (the products model's kernels)
        cfg = dict(cfg)
        cfg["outdims"] = list(cfg["outdims"][:-1]) + [int(sys.argv[3])]
    Ws, As, Wo = datasets.init_params(cfg["heads"], cfg["outdims"], cfg["I"], cfg["C"], 5)
    p = dict(row_ptr=ds["row_ptr"], col_idx=ds["col_idx"], X=ds["X"], labels=ds["labels"], Ws=Ws, As=As, Wo=Wo,
             heads=cfg["heads"], outdims=cfg["outdims"], C=cfg["C"])
eng = make_engine(gatx, p, gemm_mode=1, keep_debug=True)
if which in ("race", "race0"):  # the same backward several times: a race shows as a changing set of wrong edges
    l = len(p["heads"]) - 1
    H = p["heads"][l]
    E = len(p["col_idx"])
    if which == "race":  # with one oracle pass beside it
        ref = make_oracle(orc, p)
        ref.forward(); ref.backward()
        ga_ref = ref.tensor(orc.T_GA, l).ravel()
    else:
        ga_ref = np.zeros(p["heads"][l] * p["outdims"][l], np.float32)
    prev = None
    for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
        eng.forward(); eng.backward()
        ge = eng.tensor(gatx.T_GE, l).reshape(E, H).copy()
        ga = eng.tensor(gatx.T_GA, l).copy()
        gw = eng.tensor(gatx.T_GW, l - 1).copy()
        if prev is not None:
            ch = np.nonzero(np.abs(ge - prev[0]).max(axis=1) > 0)[0]
            print("run %d: ga vs oracle %.3e; edges whose ge changed since the last run: %d %s; ga bitwise equal %s; gW(l-1) equal %s" %
                  (it, rel_err(ga, ga_ref), len(ch), (ch[:12] % 256).tolist(), np.array_equal(ga, prev[1]), np.array_equal(gw, prev[2])), flush=True)
        else:
            print("run 0: ga vs oracle %.3e" % rel_err(ga, ga_ref), flush=True)
        prev = (ge, ga, gw)
    sys.exit(0)
ref = make_oracle(orc, p)
eng.forward(); eng.backward()
ref.forward(); ref.backward()
rp, ci = np.asarray(p["row_ptr"], np.int64), np.asarray(p["col_idx"], np.int64)
E = len(ci)
L = len(p["heads"])
X = p["X"]
for l in range(L):
    H, D = p["heads"][l], p["outdims"][l]
    F = H * D
    o = orc.layer_backward(p["row_ptr"], p["col_idx"], H, D, X, ref.tensor(orc.T_W, l), ref.tensor(orc.T_A, l),
                           ref.tensor(orc.T_PL, l), ref.tensor(orc.T_PR, l), ref.tensor(orc.T_ALPHA, l), ref.tensor(orc.T_GH, l))
    print("layer", l, "H", H, "D", D)
    for name, a, b in (("Hout", eng.tensor(gatx.T_HOUT, l), ref.tensor(orc.T_HOUT, l).ravel()),
                       ("g_h", eng.tensor(gatx.T_GH, l), ref.tensor(orc.T_GH, l).ravel()),
                       ("alpha", eng.tensor(gatx.T_ALPHA, l).reshape(E, H).T, ref.tensor(orc.T_ALPHA, l)),
                       ("ge", eng.tensor(gatx.T_GE, l).reshape(E, H).T, o["ge"]),
                       ("gPl", eng.tensor(gatx.T_GPL, l), o["gPl"].ravel()),
                       ("gPr", eng.tensor(gatx.T_GPR, l), o["gPr"].ravel()),
                       ("gW", eng.tensor(gatx.T_GW, l), ref.tensor(orc.T_GW, l).ravel()),
                       ("ga", eng.tensor(gatx.T_GA, l), ref.tensor(orc.T_GA, l).ravel())):
        print("   %-6s rel_err %.3e" % (name, rel_err(a, b)), flush=True)
    if l == L - 1:
        ge = eng.tensor(gatx.T_GE, l).reshape(E, H)
        d = np.abs(ge - o["ge"].T).max(axis=1)
        w = np.argsort(-d)[:8]
        dst = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        print("   worst ge edges", [(int(e), int(e) % 256, int(dst[e]), int(rp[dst[e] + 1] - rp[dst[e]]), float(d[e])) for e in w])
        gpr = eng.tensor(gatx.T_GPR, l).reshape(-1, F)
        dr = np.abs(gpr - o["gPr"]).max(axis=1)
        w = np.argsort(-dr)[:8]
        print("   worst gPr rows", [(int(r), int(rp[r]), int(rp[r + 1] - rp[r]), float(dr[r]), float(np.abs(o["gPr"][r]).max())) for r in w])
        gh_e, gh_r = eng.tensor(gatx.T_GH, l).reshape(-1, F), ref.tensor(orc.T_GH, l).reshape(-1, F)
        dg = np.abs(gh_e - gh_r).max(axis=1)
        print("   rows with g_h error > 1e-4 of max:", int((dg > 1e-4 * np.abs(gh_r).max()).sum()), "of", len(dg))
        for r in np.argsort(-dg)[:6]:
            k = int(np.argmax(np.abs(gh_e[r] - gh_r[r])))
            print("   g_h row %d start %d (mod 256 = %d) deg %d err %.3e at col %d: eng %s ref %s Hout %s" %
                  (r, rp[r], rp[r] % 256, rp[r + 1] - rp[r], dg[r], k, gh_e[r, max(0, k - 2):k + 3], gh_r[r, max(0, k - 2):k + 3],
                   eng.tensor(gatx.T_HOUT, l).reshape(-1, D)[r, max(0, k - 2):k + 3]))
        bad = np.nonzero(dg > 1e-4 * np.abs(gh_r).max())[0]
        print("   bad rows (first 40):", bad[:40].tolist())
        print("   bad rows start mod 256 (first 40):", (rp[bad[:40]] % 256).tolist())
        # ga recomputed on the host from the ENGINE's ge, P_l, P_r: is the kernel's accumulation or its input off?
        Pl, Pr = eng.tensor(gatx.T_PL, l).reshape(-1, F), eng.tensor(gatx.T_PR, l).reshape(-1, F)
        ga = np.zeros(F, np.float64)
        for lo in range(0, E, 200000):
            hi = min(E, lo + 200000)
            s = Pl[ci[lo:hi]].astype(np.float64) + Pr[dst[lo:hi]]
            u = np.where(s > 0, 1.0, 0.01) * np.repeat(ge[lo:hi].astype(np.float64), D, axis=1)
            ga += (u * s).sum(axis=0)
        print("   ga engine vs host sum over engine ge: %.3e; host sum vs oracle: %.3e" %
              (rel_err(eng.tensor(gatx.T_GA, l), ga), rel_err(ga, ref.tensor(orc.T_GA, l).ravel())))
    X = ref.tensor(orc.T_HOUT, l)
eng.close()
