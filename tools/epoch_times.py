"""Per-epoch device times on the products shape: whole-epoch phase timer vs the stopwatch bench.py uses."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-attention-network-gatv2-_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import bench, datasets, gatx
ds = bench.load_workload("products", 1.0, 0, 1, lambda: None)
cfg = ds["cfg"]
opt = os.environ.get("OPT", cfg["optimizer"])
lr = float(os.environ.get("LR", cfg["lr"]))
print("optimizer", opt, "lr", lr)
eng = gatx.Engine(cfg["heads"], cfg["outdims"], optimizer=opt, lr=lr, clip=cfg["clip"])
eng.set_graph(np.asarray(ds["row_ptr"]), np.asarray(ds["col_idx"]))
eng.set_features(np.asarray(ds["X"]))
eng.set_labels(np.asarray(ds["labels"]), cfg["C"])
eng.init_params(1234)
for t in range(1, 4):
    eng.train_epoch(t, want_loss=False)
eng.sync()
eng.enable_timing(True)
for t in range(4, 4 + int(os.environ.get('EPOCHS', '8'))):
    eng.timer_start()
    t0 = time.perf_counter()
    eng.train_epoch(t, want_loss=False)
    host_ms = (time.perf_counter() - t0) * 1e3
    sw = eng.timer_stop()
    ph = eng.timing()
    print("epoch %2d: stopwatch %.2f ms, edge_fwd %.2f edge_bwd %.2f gemm %.2f, loss %.4f" %
          (t, sw, ph["edge_fwd"], ph["edge_bwd"], ph["gemm_fwd"] + ph["gemm_bwd"], eng.loss_acc()[0]))
eng.timer_start()
for t in range(100, 105):
    eng.train_epoch(t, want_loss=False)
print("5 epochs back to back: %.2f ms each" % (eng.timer_stop() / 5))
