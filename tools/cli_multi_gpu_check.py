import os, re, subprocess, sys, tempfile
sys.path.insert(0, "graph-attention-network-gatv2-_b200")
import build as gatx_build, datasets
cli = gatx_build.build_cli()
tmp = tempfile.mkdtemp()
ds = datasets.make_dataset("arxiv", 0.2)
datasets.write_txt(os.path.join(tmp, "g"), ds)
base = [cli, "--num-layers", "3", "--heads", "4,4,1", "--outdims", "64,64,64", "--epochs", "6", "--optimizer", "adam",
        "--lr", "0.01", "--dataset", "g", "--data-root", tmp, "--seed", "3", "--gemm", "fp32"]
curves = {}
for g in ("1", "8"):
    r = subprocess.run(base + ["--gpus", g], capture_output=True, text=True, timeout=300)
    print("gpus", g, "rc", r.returncode, r.stderr[-300:])
    curves[g] = [float(x) for x in re.findall(r"Avg Loss: ([0-9.]+)", r.stdout)]
print(curves)
ok = len(curves["8"]) == 6 and all(abs(a - b) < 2e-4 * max(1.0, a) for a, b in zip(curves["1"], curves["8"]))
print("CLI 8-GPU matches 1-GPU:", ok)
