"""Epoch time of the small BASELINE configs with eager launches vs CUDA-graph replay (gatx_set_cuda_graph):
python tools/graph_replay_times.py [sample cora pubmed arxiv] -> one JSON line per workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-attention-network-gatv2-_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import bench, gatx

for name in (sys.argv[1:] or ["sample", "cora", "pubmed", "arxiv"]):
    ds = bench.load_workload(name, 1.0, 0, 1, lambda: None)
    cfg = ds["cfg"]
    out = {"workload": name, "N": cfg["N"], "E": int(np.asarray(ds["row_ptr"])[-1])}
    losses = {}
    for mode, label in ((0, "eager"), (1, "graph")):
        eng = gatx.Engine(cfg["heads"], cfg["outdims"], optimizer=cfg["optimizer"], lr=cfg["lr"], clip=cfg["clip"])
        eng.set_graph(np.asarray(ds["row_ptr"]), np.asarray(ds["col_idx"]))
        eng.set_features(np.asarray(ds["X"]))
        eng.set_labels(np.asarray(ds["labels"]), cfg["C"])
        eng.init_params(1234)
        eng.set_cuda_graph(mode)
        for t in range(1, 6):
            eng.train_epoch(t, want_loss=False)
        eng.sync()
        K = 200 if cfg["E"] < 2_000_000 else 40
        l0 = eng.launch_count()
        eng.timer_start()
        for t in range(6, 6 + K):
            eng.train_epoch(t, want_loss=False)
        ms = eng.timer_stop() / K
        out[label + "_ms_per_epoch"] = round(ms, 4)
        eng.timer_start()
        for t in range(6 + K, 6 + 2 * K):
            eng.train_epoch(t)  # loss / accuracy read back every epoch, as train_gatx prints them
        out[label + "_ms_per_epoch_with_readback"] = round(eng.timer_stop() / K, 4)
        out[label + "_launches_per_epoch"] = (eng.launch_count() - l0) // K
        assert eng.cuda_graph_active() == bool(mode)
        losses[label] = eng.loss_acc()
        eng.close()
    assert losses["eager"] == losses["graph"], losses  # bit-identical training
    out["speedup"] = round(out["eager_ms_per_epoch"] / out["graph_ms_per_epoch"], 2)
    out["loss_after"] = losses["graph"][0]
    print(json.dumps(out), flush=True)
