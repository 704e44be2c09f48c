"""Where do the parameters of a 2-GPU run start to differ from the 1-GPU run?  (python tools/multi_gpu_weight_drift.py, on a
box with >= 2 GPUs.)  train_gatx --dump-weights after k epochs for the transports / block counts of the exchange."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-attention-network-gatv2-_b200"))
import datasets
cli = os.path.join(ROOT, "graph-attention-network-gatv2-_b200", "train_gatx")
tmp = tempfile.mkdtemp()
datasets.write_txt(os.path.join(tmp, "g"), datasets.make_dataset("arxiv", 0.05))
base = [cli, "--num-layers", "3", "--heads", "4,4,1", "--outdims", "32,32,128", "--optimizer", "adam", "--lr", "0.01", "--clip",
        "--dataset", "g", "--data-root", tmp, "--seed", "3", "--gemm", "fp32"]
def run(gpus, env, epochs):
    d = tempfile.mkdtemp()
    e = dict(os.environ); e.update(env)
    r = subprocess.run(base + ["--gpus", str(gpus), "--epochs", str(epochs), "--dump-weights", d], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    return np.fromfile(os.path.join(d, "W.bin"), np.float32), r.stdout.count("Avg Loss")
for epochs in (1, 2, 4, 6):
    ref, _ = run(1, {}, epochs)
    for name, env in (("2gpu bulk K=3", {"GATX_HALO_BLOCKS": "3"}), ("2gpu bulk K=1", {"GATX_HALO_BLOCKS": "1"}),
                      ("2gpu sm K=1", {"GATX_HALO_BLOCKS": "1", "GATX_HALO_MODE": "sm"}), ("2gpu nccl", {"GATX_NO_P2P": "1"})):
        w, _ = run(2, env, epochs)
        d = np.abs(w - ref)
        print("epochs %d %-14s max|dW| %.3e  mean|dW| %.3e  frac>1e-4 %.4f  (max|W| %.3f)" % (epochs, name, d.max(), d.mean(), np.mean(d > 1e-4), np.abs(ref).max()), flush=True)
