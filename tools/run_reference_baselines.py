"""GPU box: time the reference's edge- and node-based binaries (oracle/_ref, rebuilt for sm_100a, otherwise
unmodified) and train_gatx on the same text datasets with the same flags.  Writes gpurun_out/ref_baselines.json.
Median of the printed per-epoch ` total time:` over epochs 2..k (BASELINE.md timing protocol)."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-attention-network-gatv2-_b200"))
import datasets  # noqa: E402

CASES = [("sample", 1.0, 6), ("cora", 1.0, 6), ("pubmed", 1.0, 5), ("arxiv", 1.0, 3), ("products", 0.001, 3)]
BINS = {"reference_edge": os.path.join(ROOT, "oracle", "_ref", "edge_ref"),
        "reference_node": os.path.join(ROOT, "oracle", "_ref", "node_ref"),
        "gatx": os.path.join(ROOT, "graph-attention-network-gatv2-_b200", "train_gatx")}
cap = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
out = []
for name, scale, epochs in CASES:
    ds = datasets.make_dataset(name, scale)
    cfg = ds["cfg"]
    tmp = tempfile.mkdtemp(prefix="gatx_base_")
    datasets.write_txt(os.path.join(tmp, name), ds)
    flags = ["--num-layers", str(len(cfg["heads"])), "--heads", ",".join(map(str, cfg["heads"])), "--outdims",
             ",".join(map(str, cfg["outdims"])), "--optimizer", cfg["optimizer"], "--lr", str(cfg["lr"]), "--dataset", name,
             "--data-root", tmp] + (["--clip"] if cfg["clip"] else [])
    for impl, binp in BINS.items():
        ep = epochs if impl != "gatx" else max(epochs, 12)
        rec = dict(workload=name, N=cfg["N"], E=cfg["E"], impl=impl, flags=" ".join(flags[:8]), epochs=ep)
        t0 = time.time()
        try:
            r = subprocess.run([binp] + flags + ["--epochs", str(ep)] + (["--seed", "1"] if impl == "gatx" else []),
                               capture_output=True, text=True, timeout=cap)
            times = [float(x) for x in re.findall(r"total time: ([0-9.eE+-]+) ms", r.stdout)]
            losses = [float(x) for x in re.findall(r"Avg Loss: ([0-9.eE+-]+|nan|inf|-nan)", r.stdout)]
            rec.update(rc=r.returncode, epoch_ms_median=float(np.median(times[1:])) if len(times) > 1 else None,
                       epoch_ms_first=times[0] if times else None, first_loss=losses[0] if losses else None,
                       wall_s=time.time() - t0)
            if rec["epoch_ms_median"]:
                rec["edges_per_s"] = cfg["E"] / (rec["epoch_ms_median"] * 1e-3)
        except subprocess.TimeoutExpired as e:
            so = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            times = [float(x) for x in re.findall(r"total time: ([0-9.eE+-]+) ms", so)]
            rec.update(rc=None, timeout_s=cap, epochs_finished=len(times), epoch_ms_first=times[0] if times else None,
                       note="did not finish %d epochs within the %.0f s cap" % (ep, cap))
        out.append(rec)
        print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_baselines.json"), "w"), indent=1)
