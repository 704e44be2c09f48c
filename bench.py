#!/usr/bin/env python
"""bench.py -- GATv2 full-batch training throughput on synthetic graphs of the reference's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload products] [--impl gatx|reference]

A "step" is one training epoch (forward + loss/accuracy + backward + clip/optimizer + gradient
reset), the span the reference times (GATv2_edge_based.cu:1371 -> 1639).  The metric is
edges/s = E / epoch time (BASELINE.json); epochs/s is reported beside it.  N > 1 is launched by
torchrun (one rank per GPU); destination rows are partitioned across ranks and the total work is
fixed ("strong" scaling: the same graph is trained by more GPUs).

--impl reference runs the reference's own edge-based CUDA binary (oracle/_ref/edge_ref, built from
/root/reference by oracle/Makefile) on a bounded sample of the same workload written in the
reference's text format -- the reference has no CPU path; if the binary is absent the CPU oracle
port is timed instead.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "graph-attention-network-gatv2-_b200")
sys.path.insert(0, PKG)

import datasets  # noqa: E402


# ------------------------------------------------------------------ helpers
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


def load_workload(name, scale, rank, world, barrier):
    """Rank 0 generates the synthetic dataset once and caches it under /dev/shm; the others map it."""
    cache = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                         "gatx_%s_%g" % (name, scale))
    done = os.path.join(cache, "done")
    if rank == 0 and not os.path.exists(done):
        os.makedirs(cache, exist_ok=True)
        ds = datasets.make_dataset(name, scale)
        for k in ("row_ptr", "col_idx", "X", "labels"):
            np.save(os.path.join(cache, k + ".npy"), ds[k])
        open(done, "w").write("ok")
    barrier()
    while not os.path.exists(done):
        time.sleep(0.5)
    ds = {k: np.load(os.path.join(cache, k + ".npy"), mmap_mode="r") for k in ("row_ptr", "col_idx", "X", "labels")}
    cfg = dict(datasets.CONFIGS[name])
    cfg.update(N=len(ds["labels"]), E=len(ds["col_idx"]))
    ds["cfg"] = cfg
    return ds


def flags_of(cfg):
    return "--num-layers %d --heads %s --outdims %s --optimizer %s%s" % (
        len(cfg["heads"]), ",".join(map(str, cfg["heads"])), ",".join(map(str, cfg["outdims"])), cfg["optimizer"],
        " --clip" if cfg["clip"] else "")


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(name, budget_s=20.0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    cfgf = datasets.CONFIGS[name]
    # bounded sample: same feature / class / model dims, N and E scaled down together
    scale = 1.0 if cfgf["N"] <= 4096 else max(2048.0 / cfgf["N"], 0.001)
    ds = datasets.make_dataset(name, scale)
    cfg = ds["cfg"]
    Ws, As, Wo = datasets.init_params(cfg["heads"], cfg["outdims"], cfg["I"], cfg["C"], 0)
    m = orc.Model(cfg["heads"], cfg["outdims"], ds["row_ptr"], ds["col_idx"], ds["X"], ds["labels"],
                  num_classes=cfg["C"], optimizer=cfg["optimizer"], clip=cfg["clip"], lr=cfg["lr"])
    for l in range(len(Ws)):
        m.set_params(l, Ws[l], As[l])
    m.set_wo(Wo)
    m.epoch(1)
    t0, n = time.perf_counter(), 0
    while True:
        m.epoch(2 + n)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 5:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": cfg["E"] / dt, "unit": "edges/s", "cores": orc.num_threads(), "kind": "port",
            "sample": "%s shape scaled to N=%d E=%d (same feats/classes/model), %d epochs of the OpenMP C oracle, "
                      "%.2f s/epoch" % (name, cfg["N"], cfg["E"], n, dt)}


# ------------------------------------------------------------------ the reference's own binary
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "edge_ref")


def time_reference_binary(name, ds, epochs, warmup, timeout):
    """Runs oracle/_ref/edge_ref (GATv2_edge_based.cu rebuilt for sm_100a, unmodified but for its first line) on `ds`
    written in the reference's text format; returns (ms per epoch over the epochs after `warmup`, wall seconds).
    The printed ' total time:' is the reference's own timer (EB:1371 -> 1639)."""
    cfg = ds["cfg"]
    tmp = tempfile.mkdtemp(prefix="gatx_ref_")
    datasets.write_txt(os.path.join(tmp, name), ds)
    cmd = [REF_BIN] + flags_of(cfg).split() + ["--epochs", str(epochs), "--lr", str(cfg["lr"]), "--dataset", name,
                                               "--data-root", tmp]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    times = [float(x) for x in re.findall(r"total time: ([0-9.eE+-]+) ms", out.stdout)]
    if out.returncode != 0 or len(times) < epochs:
        raise RuntimeError("reference binary failed (rc=%d, %d epochs parsed)" % (out.returncode, len(times)))
    return float(np.mean(times[warmup:])), wall


def time_gatx_small(name, ds, epochs, warmup, gemm_mode):
    """Our arm of the same-config check: the same dataset and flags through the C ABI, ms per epoch (CUDA events)."""
    import gatx
    cfg = ds["cfg"]
    eng = gatx.Engine(cfg["heads"], cfg["outdims"], optimizer=cfg["optimizer"], lr=cfg["lr"], clip=cfg["clip"],
                      gemm_mode=gemm_mode)
    eng.set_graph(ds["row_ptr"], ds["col_idx"])
    eng.set_features(ds["X"])
    eng.set_labels(ds["labels"], cfg["C"])
    eng.init_params(1234)
    t = 0
    for _ in range(warmup):
        t += 1
        eng.train_epoch(t, want_loss=False)
    eng.sync()
    eng.timer_start()
    for _ in range(epochs):
        t += 1
        eng.train_epoch(t, want_loss=False)
    ms = eng.timer_stop() / epochs
    # with the loss / accuracy read back every epoch, as the reference prints them (wall clock around the calls)
    t0 = time.perf_counter()
    for _ in range(epochs):
        t += 1
        eng.train_epoch(t, want_loss=True)
    eng.sync()
    ms_sync = (time.perf_counter() - t0) / epochs * 1e3
    eng.close()
    return ms, ms_sync


def same_config_check(args):
    """Both arms on ONE identical graph and ONE identical flag set, in the same run on the same GPU: the full
    arxiv-shaped config (BASELINE config 4) and the pubmed-shaped config (config 3).  The headline products shape
    cannot be run by the reference within hours (its per-edge mat-vec recompute and serialised shared-memory atomics,
    EB:698-798), so this block is the like-for-like speed-up a reader may quote."""
    import gatx
    out = {}
    if not os.path.exists(REF_BIN):
        return {"unavailable": "oracle/_ref/edge_ref not built"}
    for name, ref_epochs in (("pubmed", 6), ("arxiv", 3)):
        try:
            ds = datasets.make_dataset(name)
            cfg = ds["cfg"]
            ref_ms, wall = time_reference_binary(name, ds, ref_epochs + 1, 1, args.ref_timeout)
            ms, ms_sync = time_gatx_small(name, ds, 20, 3, gatx.GEMM_FP32_SIMT if args.fp32 else gatx.GEMM_TF32_TC)
            out[name] = {"config": "%s-shaped N=%d E=%d feats=%d classes=%d, %s, lr %g (identical files and flags for "
                                   "both arms)" % (name, cfg["N"], cfg["E"], cfg["I"], cfg["C"], flags_of(cfg), cfg["lr"]),
                         "reference_ms_per_epoch": ref_ms, "reference_epochs_timed": ref_epochs,
                         "reference_timer": "the binary's own ' total time:' line (EB:1371->1639), 1 warm-up epoch",
                         "gatx_ms_per_epoch": ms, "gatx_ms_per_epoch_with_loss_readback": ms_sync,
                         "gatx_epochs_timed": 20, "speedup": ref_ms / ms, "speedup_with_loss_readback": ref_ms / ms_sync,
                         "reference_edges_per_s": cfg["E"] / (ref_ms * 1e-3), "gatx_edges_per_s": cfg["E"] / (ms * 1e-3),
                         "reference_wall_s": wall}
        except Exception as e:  # a reference failure must not lose the benchmark line
            out[name] = {"unavailable": str(e)[:200]}
    return out


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    cfgf = datasets.CONFIGS[name]
    binp = REF_BIN
    line = {"impl": "reference", "metric": "train_edges_per_s", "unit": "edges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if os.path.exists(binp) and not args.cpu_reference:
        # the reference's own CUDA binary (it has no CPU path), single GPU, bounded sample
        scale = args.ref_scale if args.ref_scale else (1.0 if cfgf["N"] <= 20000 else max(2000.0 / cfgf["N"], 0.0005))
        ds = datasets.make_dataset(name, scale)
        cfg = ds["cfg"]
        epochs = args.steps + args.warmup
        try:
            ms, wall = time_reference_binary(name, ds, epochs, args.warmup, args.ref_timeout)
        except Exception as e:
            line.update({"unavailable": str(e)[:200]})
            print(json.dumps(line))
            return
        val = cfg["E"] / (ms * 1e-3)
        sample = ("reference edge-based CUDA binary (GATv2_edge_based.cu rebuilt for sm_100a) on 1 GPU, %s shape scaled "
                  "to N=%d E=%d, flags '%s', mean of printed epoch times after %d warm-up epochs (wall %.1f s)"
                  % (name, cfg["N"], cfg["E"], flags_of(cfg), args.warmup, wall))
        line.update({"value": val, "ms_per_step": ms, "epochs_per_s": 1e3 / ms,
                     "config": {"workload": "%s-shaped sample N=%d E=%d %s" % (name, cfg["N"], cfg["E"], flags_of(cfg)),
                                "sample_of": "%s (N=%d E=%d)" % (name, cfgf["N"], cfgf["E"]),
                                "is_sample": scale != 1.0,
                                "note": "this arm is a BOUNDED SAMPLE of the workload (the reference needs hours per "
                                        "epoch on the full products shape); the like-for-like comparison of both arms "
                                        "on identical full-size graphs (arxiv, pubmed) is the same_config_check block "
                                        "of the gatx arm's line",
                                "l2": "sample smaller than L2 (the reference binary has no flush hook)"},
                     "cpu_baseline": {"value": val, "unit": "edges/s", "cores": 0, "kind": "reference", "sample": sample},
                     "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
    else:
        cb = cpu_baseline(name, budget_s=30.0)
        cb["kind"] = "port"
        line.update({"value": cb["value"], "ms_per_step": None,
                     "config": {"workload": "%s-shaped sample (CPU oracle port)" % name},
                     "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def run_gatx(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch  # plumbing only: rendezvous/barrier for N > 1 and pinned host buffers
    import gatx
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner / debug output (NCCL_DEBUG=VERSION on some boxes) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()

    ds = load_workload(args.workload, args.scale, rank, world, barrier)
    cfg = ds["cfg"]
    N, E = cfg["N"], cfg["E"]
    # Learning rate: the reference back-propagates the SUMMED loss (EB:572), so its default SGD step of 1e-4 is an
    # effective 245 per node on 2.45 M nodes and the run diverges to the loss clamp -log(1e-12) within five epochs
    # (DESIGN.md D9).  On those saturated values the power-capped GPU clocks ~10 % higher and every kernel runs faster
    # (measured: 178 ms/epoch against 199 ms on finite, slowly converging values), so the benchmark trains with a step
    # that keeps the values healthy: 4e-8 x N = 0.1 per node.  --lr restores any other value.
    if args.lr is not None:
        cfg["lr"] = args.lr
    elif args.workload == "products" and cfg["optimizer"] == "sgd":
        cfg["lr"] = 4e-8
    eng = gatx.Engine(cfg["heads"], cfg["outdims"], optimizer=cfg["optimizer"], lr=cfg["lr"], clip=cfg["clip"],
                      device=local_rank, rank=rank, world=world,
                      gemm_mode=gatx.GEMM_FP32_SIMT if args.fp32 else (gatx.GEMM_3XTF32_TC if args.x3 else gatx.GEMM_TF32_TC))
    if world > 1:
        ids = [gatx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        # NCCL prints its version banner with printf when the box sets NCCL_DEBUG=VERSION (NCCL_DEBUG_FILE does not move
        # it): file descriptor 1 points at stderr while the communicator is created, stdout keeps to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            eng.comm_init(ids[0])
        finally:
            import ctypes
            ctypes.CDLL(None).fflush(None)  # the banner sits in C stdio's buffer when stdout is a pipe
            os.dup2(saved, 1)
            os.close(saved)
    # pinned host copies of the per-epoch inputs for the end-to-end leg
    Xp = torch.empty((N, cfg["I"]), dtype=torch.float32, pin_memory=True)
    Xp.numpy()[:] = ds["X"]
    yp = torch.empty((N,), dtype=torch.int32, pin_memory=True)
    yp.numpy()[:] = ds["labels"]
    eng.set_graph(np.asarray(ds["row_ptr"]), np.asarray(ds["col_idx"]))
    eng.set_features_ptr(Xp.data_ptr(), cfg["I"])
    eng.set_labels_ptr(yp.data_ptr(), cfg["C"])
    eng.init_params(1234)
    info = eng.graph_info()
    halo = None
    if world > 1:
        # NVLink peer-memory halo exchange: every rank publishes the IPC handles of its P_l / gP_l buffers
        push_rows = eng.halo_rows()
        if not args.no_p2p:
            blobs = [None] * world
            dist.all_gather_object(blobs, eng.peer_export())
            eng.peer_import(blobs)
        tmode = os.environ.get("GATX_HALO_MODE", "pull")
        halo = {"exchange": ("NVLink peer memory, pipelined over row blocks under the edge passes, flag barriers in peer "
                             "memory; transport: " +
                             ("bulk-copy (TMA) push + scatter kernels (global -> shared ring -> peer global), staging + local "
                              "ordered sum" if tmode == "bulk" else "ld / st kernels: owner pushes P_l rows (forward), owner "
                              "pulls and sums partial gP_l rows in rank order (backward); halo rows only"))
                if eng.halo_active() else "nccl broadcast/reduce",
                "rows_pushed_per_layer_rank0": push_rows,
                "allgather_rows_rank0": (world - 1) * (info["row_end"] - info["row_begin"])}
    t = 0
    for _ in range(args.warmup):
        t += 1
        eng.train_epoch(t, want_loss=False)
    eng.sync()
    launches0 = eng.launch_count()
    eng.enable_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    eng.sync()
    phase_acc = {}
    eng.timer_start()
    for _ in range(args.steps):
        t += 1
        eng.train_epoch(t, want_loss=False)
        if args.phase_times:
            for k, v in eng.timing().items():
                phase_acc[k] = phase_acc.get(k, 0.0) + v
    total_ms = eng.timer_stop()
    eng.sync()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.summary()
    launches = eng.launch_count() - launches0
    if not args.phase_times:
        # per-phase device times of the last timed epoch (events were recorded inside the timed region)
        phase_acc = {k: v * args.steps for k, v in eng.timing().items()}
    kernel_ms = [eng.edge_kernel_ms(l) for l in range(len(cfg["heads"]))]  # per-launch CUDA-event times, last epoch
    halo_stats = eng.halo_stats() if world > 1 else None
    eng.enable_timing(False)
    loss, acc = eng.loss_acc()
    if dist is not None:
        tt = torch.tensor([total_ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms = total_ms / args.steps

    # end-to-end: host buffers in, scalars out, every step (same public API a user calls)
    e2e_steps = max(1, args.steps)
    barrier()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t += 1
        eng.set_features_ptr(Xp.data_ptr(), cfg["I"])
        eng.set_labels_ptr(yp.data_ptr(), cfg["C"])
        l2, a2 = eng.train_epoch(t, want_loss=True)
    eng.sync()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    rows = info["row_end"] - info["row_begin"]
    # a rank copies its own feature rows and labels; the row blocks are all-gathered over NVLink inside set_features
    h2d = (rows if world > 1 else N) * cfg["I"] * 4 + rows * 4
    d2h = 16

    # Roofline (HBM-bound edge passes).  Algorithmic bytes per launch are SURVEY 8(d)'s formulas split by kernel
    # (b = 4 bytes per stored projected element); durations are CUDA-event times of the individual launches in
    # the last timed epoch.  The reported kernel is the one with the largest duration.
    Nl, El = rows, None
    kernels = []
    fb = bb = 0.0
    for l in range(len(cfg["heads"])):
        f, b = eng.edge_bytes(l)
        fb += f
        bb += b
        H, D = cfg["heads"][l], cfg["outdims"][l]
        F = H * D
        El = (f - 4.0 * (Nl + 1) - Nl * F * 8.0) / (4.0 + 4.0 * F + 4.0 * H)  # local edges from the fwd formula
        p1 = 4.0 * (Nl + 1) + El * (4.0 + 4.0 * F) + 8.0 * H * El + 12.0 * Nl * F
        p2 = b - p1
        ms3 = kernel_ms[l]
        # one head of 128 / 64 floats runs the two-edges-per-trip kernels of edge_stream_pair.inc (ncu names *_pair_kernel)
        fam = "pair" if (H == 1 and (D == 64 or (D == 128 and not os.environ.get("GATX_NO_PAIR")))) else "stream"
        for name, key, by in (("edge_fwd_%s_kernel" % fam, "fwd", f), ("edge_bwd_dst_%s_kernel" % fam, "bwd_dst", p1),
                              ("edge_bwd_src_%s_kernel" % fam, "bwd_src", p2)):
            if ms3[key] > 0:
                kernels.append({"kernel": name, "layer": l, "F": F, "ms": ms3[key], "algorithmic_gb": by / 1e9,
                                "gbs": by / 1e9 / (ms3[key] * 1e-3)})
    edge_ms = (phase_acc.get("edge_fwd", 0.0) + phase_acc.get("edge_bwd", 0.0)) / args.steps
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    agg = (fb + bb) / 1e9 / (edge_ms * 1e-3) if edge_ms > 0 else 0.0
    traffic = None
    tj = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if kernels:
        dom = max(kernels, key=lambda k: k["ms"])
        if os.path.exists(tj) and world == 1:
            t = json.load(open(tj)).get("%s:%s:F%d" % (args.workload, dom["kernel"], dom["F"]))
            traffic = t["dram_bytes_per_launch"] / 1e9 if t and abs(t["scale"] - args.scale) < 1e-9 else None
    else:
        dom = {"kernel": "edge kernels (narrow-row path, phase timing only)", "layer": -1, "F": 0, "ms": edge_ms,
               "algorithmic_gb": (fb + bb) / 1e9, "gbs": agg}
    roof = {"bound": "hbm", "kernel": "%s (layer %d, %d-float rows)" % (dom["kernel"], dom["layer"], dom["F"]),
            "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak,
            "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
            "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write, profiles/)",
            "algorithmic_gb_per_launch": dom["algorithmic_gb"], "ms_per_launch": dom["ms"],
            # frac can exceed 1: the peak is a COPY bandwidth (read + write streams) and part of the gathers is served
            # by L2 (hot rows are fetched evict_last), so DRAM traffic is below the algorithmic bytes
            "dram_gbs_from_traffic": (traffic / (dom["ms"] * 1e-3)) if traffic else None,
            "all_edge_kernels": kernels,
            "edge_passes_aggregate": {"algorithmic_gb_per_epoch": (fb + bb) / 1e9, "ms_per_epoch": edge_ms,
                                      "gbs": agg, "frac": agg / peak}}
    if halo is not None and halo_stats is not None and halo_stats["push_ms"] > 0:
        # rank 0's exchange kernels in the last timed epoch: bytes over NVLink / time the kernels were running (they
        # run underneath the edge passes; the EXPOSED part is phase_ms_per_epoch.comm)
        halo["nvlink_fwd_gbs_rank0"] = halo_stats["push_bytes"] / 1e9 / (halo_stats["push_ms"] * 1e-3)
        halo["nvlink_bwd_gbs_rank0"] = (halo_stats["pull_bytes"] / 1e9 / (halo_stats["pull_ms"] * 1e-3)
                                        if halo_stats["pull_ms"] > 0 else None)
        halo["fwd_gb_per_epoch_rank0"] = halo_stats["push_bytes"] / 1e9
        halo["bwd_gb_per_epoch_rank0"] = halo_stats["pull_bytes"] / 1e9
        halo["fwd_busy_ms_rank0"] = halo_stats["push_ms"]
        halo["bwd_busy_ms_rank0"] = halo_stats["pull_ms"]
    if rank == 0:
        line = {
            "metric": "train_edges_per_s", "value": E / (ms * 1e-3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "epochs_per_s": 1e3 / ms,
            # every (edge, head) pair is traversed once per layer in the forward and twice in the backward
            "edge_head_traversals_per_s": E * sum(cfg["heads"]) / (ms * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else ("f32 (3xTF32 tensor-core projections)" if args.x3 else "f32 (tf32 tensor-core projections)"), "data": "synthetic",
            "config": {"workload": "%s-shaped synthetic graph N=%d E=%d feats=%d classes=%d, %s, lr %g"
                                   % (args.workload, N, E, cfg["I"], cfg["C"], flags_of(cfg), cfg["lr"]),
                       "max_in_degree": info["max_degree"], "parallelism": "dst-row partition x%d" % world,
                       "halo": halo,
                       "l2": "inputs exceed L2 (per-epoch working set >> 126 MB)" if E * 4 * 64 > 126e6 else
                             "working set may fit L2 (small workload)"},
            "e2e": {"value": E * e2e_steps / e2e_s, "unit": "edges/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / e2e_steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "phase_ms_per_epoch": {k: v / args.steps for k, v in phase_acc.items()},
            "final_loss": loss, "final_acc": acc,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        if world == 1 and not args.no_same_config:
            line["same_config_check"] = same_config_check(args)
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gatx", choices=["gatx", "reference"])
    ap.add_argument("--lr", type=float, default=None, help="learning rate (default: the workload's, see run_gatx)")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: NCCL collectives instead of the peer-memory halo kernels")
    ap.add_argument("--workload", default="products", choices=list(datasets.CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--fp32", action="store_true", help="fp32 CUDA-core GEMMs instead of TF32 tensor cores")
    ap.add_argument("--x3", action="store_true", help="3xTF32: fp32-grade GEMMs on the tensor cores (hi/lo operand split)")
    ap.add_argument("--phase-times", action="store_true", help="sync after every epoch to sum per-phase times")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-same-config", action="store_true",
                    help="skip the same-config comparison with the reference binary (arxiv + pubmed, ~2 min)")
    ap.add_argument("--cpu-reference", action="store_true", help="--impl reference: time the CPU oracle port")
    ap.add_argument("--ref-scale", type=float, default=0.0)
    ap.add_argument("--ref-timeout", type=float, default=900.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gatx(args)


if __name__ == "__main__":
    main()
