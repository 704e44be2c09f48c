"""Debug helper (GPU box): per-tensor error of TF32-TC mode and fp32-SIMT mode against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("graph-attention-network-gatv2-_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import gatx, orc
from helpers import make_engine, make_oracle, make_problem, rel_err

N, E, I, C, heads, outdims, kind, hub = (300, 2000, 33, 5, (4, 4, 1), (64, 64, 64), "rmat", None)
p = make_problem(N, E, I, C, heads, outdims, kind, seed=N, hub=hub)
ref = make_oracle(orc, p); ref.forward(); ref.backward()
engs = {m: make_engine(gatx, p, gemm_mode=m, keep_debug=True) for m in (0, 1)}
for e in engs.values():
    e.forward(); e.backward()
names = [("PL", gatx.T_PL, orc.T_PL), ("PR", gatx.T_PR, orc.T_PR), ("HOUT", gatx.T_HOUT, orc.T_HOUT),
         ("GH", gatx.T_GH, orc.T_GH), ("GW", gatx.T_GW, orc.T_GW), ("GA", gatx.T_GA, orc.T_GA)]
for l in range(len(heads)):
    for nm, tg, to in names:
        r = ref.tensor(to, l).ravel()
        t0 = engs[0].tensor(tg, l)
        l2 = np.linalg.norm(t0.astype(np.float64) - r) / np.linalg.norm(r)
        frac = np.mean(np.abs(t0 - r) > 1e-2 * np.abs(r).max())
        print(l, nm, "tc-vs-oracle max %.2e L2 %.2e frac>1%% %.2e | simt-vs-oracle %.2e  max|ref| %.3e" % (
            rel_err(t0, r), l2, frac, rel_err(engs[1].tensor(tg, l), r), np.abs(r).max()))
    for nm, tg in (("GPL", gatx.T_GPL), ("GPR", gatx.T_GPR)):
        a, b = engs[0].tensor(tg, l), engs[1].tensor(tg, l)
        print(l, nm, "tc-vs-simt %.2e max %.3e" % (rel_err(a, b), np.abs(b).max()))
# isolate the dual-operand GEMM: feed the SIMT engine's gP into numpy and compare both engines' GH of layer 0
l = 1
F, Iin = heads[l] * outdims[l], heads[l - 1] * outdims[l - 1]
W = p["Ws"][l]
for m in (0, 1):
    gPl = engs[m].tensor(gatx.T_GPL, l).reshape(N, F).astype(np.float64)
    gPr = engs[m].tensor(gatx.T_GPR, l).reshape(N, F).astype(np.float64)
    gX = gPl @ W[:, :Iin].astype(np.float64) + gPr @ W[:, Iin:].astype(np.float64)
    hout = engs[m].tensor(gatx.T_HOUT, 0).reshape(N, Iin)
    gh = gX * np.where(hout > 0, 1.0, 0.01)
    got = engs[m].tensor(gatx.T_GH, 0).reshape(N, Iin)
    d = np.abs(got - gh)
    print("mode", m, "GH0 vs numpy(gP of same engine): rel %.2e ; worst cols" % rel_err(got, gh), np.argsort(d.max(0))[-5:], "worst rows", np.argsort(d.max(1))[-5:])
