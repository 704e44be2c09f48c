"""Host-side logic of the destination-row partition, on CPU with world_size-2/3 gloo process groups.

Each rank plays one GPU with the CPU oracle as its "kernel": it owns the rows the library's
gatx_partition_rows assigns to it, projects its own rows, all-gathers P_l (the one exchange step of the
forward), runs the edge forward on its local CSR slice, and the per-rank (loss_sum, correct)
pairs are all-reduced.  The assembled result must equal the single-process oracle bit for bit.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q, halo_only=False):
    for p in (os.path.join(HERE, "..", "graph-attention-network-gatv2-_b200"), os.path.join(HERE, "..", "oracle"), HERE):
        sys.path.insert(0, p)
    import gatx
    import orc
    from helpers import make_problem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem(400, 3000, 12, 5, (4, 1), (8, 16), "rmat", seed=3, hub=150)
    rp, ci, X, y = p["row_ptr"], p["col_idx"], p["X"], p["labels"]
    N = len(rp) - 1
    bounds = gatx.partition_rows(rp, world)          # host-side C-ABI helper, no GPU needed
    assert np.array_equal(bounds, orc.partition_rows(rp, world))
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rp_loc = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    ci_loc = np.ascontiguousarray(ci[rp[r0]:rp[r1]])
    Xl = np.ascontiguousarray(X[r0:r1])
    for l, (H, D) in enumerate(zip(p["heads"], p["outdims"])):
        F = H * D
        Pl_loc, Pr_loc = orc.project(Xl, p["Ws"][l], F)
        # all-gather of P_l with unequal row counts: one broadcast per owner, like the NCCL group in libgatx
        Pl = np.zeros((N, F), np.float32)
        if halo_only:
            # halo exchange (halo_p2p.cu): an owner sends a row only to the ranks whose edge slice gathers it; rows
            # nobody here references stay poisoned, so a wrong mask shows up as NaNs in the result
            Pl[:] = np.nan
            Pl[r0:r1] = Pl_loc
            ref = [np.unique(ci[rp[bounds[q_]]:rp[bounds[q_ + 1]]]) for q_ in range(world)]
            sent = 0
            for src_rank in range(world):
                for dst_rank in range(world):
                    if src_rank == dst_rank:
                        continue
                    rows = ref[dst_rank][(ref[dst_rank] >= bounds[src_rank]) & (ref[dst_rank] < bounds[src_rank + 1])]
                    if rank == src_rank:
                        dist.send(torch.from_numpy(np.ascontiguousarray(Pl[rows])), dst=dst_rank)
                        sent += len(rows)
                    elif rank == dst_rank:
                        buf = torch.empty((len(rows), F), dtype=torch.float32)
                        dist.recv(buf, src=src_rank)
                        Pl[rows] = buf.numpy()
            assert sent == int(orc.halo_rows(rp, ci, world)[rank])  # the integer the library reports as gatx_halo_rows
        else:
            for r in range(world):
                t = torch.from_numpy(Pl[bounds[r]:bounds[r + 1]])
                if r == rank:
                    t.copy_(torch.from_numpy(Pl_loc))
                dist.broadcast(t, src=r)
        out = orc.layer_forward(rp_loc, ci_loc, H, D, Pl, Pr_loc, p["As"][l], l == len(p["heads"]) - 1)
        Xl = out["Hout"]
    z, yprob = orc.head_forward(p["Wo"], Xl)
    la = orc.loss_acc(yprob, np.ascontiguousarray(y[r0:r1]))
    red = torch.tensor([la["total"], float(la["correct"].sum())], dtype=torch.float64)
    dist.all_reduce(red)
    parts = [None] * world
    dist.all_gather_object(parts, (r0, r1, Xl, la["pred"]))
    if rank == 0:
        q.put((red.numpy(), parts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("halo_only", [False, True], ids=["allgather", "halo-only"])
@pytest.mark.parametrize("world", [2, 3])
def test_row_partition_forward_matches_single_process(world, halo_only, orc):
    from helpers import make_oracle, make_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + 10 * int(halo_only) + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, halo_only)) for r in range(world)]
    for pr in procs:
        pr.start()
    red, parts = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = make_problem(400, 3000, 12, 5, (4, 1), (8, 16), "rmat", seed=3, hub=150)
    ref = make_oracle(orc, p)
    ref.forward()
    rl = ref.loss()
    H_last = ref.tensor(orc.T_HOUT, 1)
    covered = 0
    for r0, r1, Hl, pred in sorted(parts, key=lambda t: t[0]):
        assert r0 == covered
        assert np.array_equal(Hl, H_last[r0:r1])      # same arithmetic, same order: bit-exact
        assert np.array_equal(pred, rl["pred"][r0:r1])
        covered = r1
    assert covered == 400
    assert abs(red[0] - rl["total"]) < 1e-9 * abs(rl["total"]) + 1e-9
    assert red[1] == round(rl["acc"] * 400)
