"""Destination-row partition across GPUs (NCCL inside libgatx): a 2-rank run must reproduce the 1-rank
run on the same box.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import sys
import tempfile
import time

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _problem():
    sys.path.insert(0, HERE)
    from helpers import make_problem
    return make_problem(3000, 40000, 24, 6, (4, 4, 1), (32, 32, 128), "rmat", seed=17, hub=1500)


def _worker(rank, world, idfile, q, mode):
    for p in (os.path.join(HERE, "..", "graph-attention-network-gatv2-_b200"), HERE):
        sys.path.insert(0, p)
    import gatx
    p = _problem()
    eng = gatx.Engine(p["heads"], p["outdims"], optimizer="adam", lr=0.01, clip=True, device=rank, rank=rank,
                      world=world, gemm_mode=mode)
    if rank == 0:
        open(idfile + ".tmp", "wb").write(gatx.comm_unique_id())
        os.rename(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.05)
    eng.comm_init(open(idfile, "rb").read())
    eng.set_graph(p["row_ptr"], p["col_idx"])
    eng.set_features(p["X"])
    eng.set_labels(p["labels"], p["C"])
    for l in range(len(p["heads"])):
        eng.set_params(l, p["Ws"][l], p["As"][l])
    eng.set_wo(p["Wo"])
    losses = [eng.train_epoch(t) for t in range(1, 5)]
    info = eng.graph_info()
    out = dict(rank=rank, losses=losses, W=[eng.tensor(gatx.T_W, l) for l in range(3)], Wo=eng.tensor(gatx.T_WO),
               rows=(info["row_begin"], info["row_end"]), pred=eng.tensor(gatx.T_PRED))
    q.put(out)
    eng.close()


@pytest.mark.parametrize("world", [2])
def test_two_ranks_match_one(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    sys.path.insert(0, os.path.join(HERE, "..", "graph-attention-network-gatv2-_b200"))
    import gatx
    from helpers import make_engine, rel_err
    mode = gatx.GEMM_FP32_SIMT
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = os.path.join(tempfile.mkdtemp(), "nccl_id")
    procs = [ctx.Process(target=_worker, args=(r, world, idfile, q, mode)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o["rank"])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = _problem()
    eng = make_engine(gatx, p, optimizer="adam", lr=0.01, clip=True, gemm_mode=mode)
    ref_losses = [eng.train_epoch(t) for t in range(1, 5)]
    for o in outs:
        for (l, a), (rl, ra) in zip(o["losses"], ref_losses):
            assert abs(l - rl) < 2e-4 * max(1.0, rl) and abs(a - ra) < 2e-3
        for l in range(3):
            assert rel_err(o["W"][l], eng.tensor(gatx.T_W, l)) < 2e-3
        assert rel_err(o["Wo"], eng.tensor(gatx.T_WO)) < 2e-3
    assert outs[0]["rows"][1] == outs[1]["rows"][0] and outs[1]["rows"][1] == 3000
    pred = np.concatenate([o["pred"] for o in outs])
    assert (pred != eng.tensor(gatx.T_PRED)).mean() < 0.01
    eng.close()
