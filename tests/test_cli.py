"""train_gatx: the reference-compatible command line (same files, flags and stdout as
GATv2_edge_based.cu:main).  Argument handling is checked on CPU; the GPU test replays the reference's own
printed loss curve from tests/golden (same dataset files, same flags, injected weights)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "graph-attention-network-gatv2-_b200")


@pytest.fixture(scope="module")
def cli():
    sys.path.insert(0, PKG)
    import build as gatx_build
    gatx_build.build()
    return gatx_build.build_cli()


def run(cli, *args, env=None):
    return subprocess.run([cli] + list(args), capture_output=True, text=True, env=env, timeout=300)


def test_argument_errors_match_reference(cli):
    r = run(cli, "--num-layers", "0")
    assert r.returncode == 1 and "Error: Number of layers must be > 0" in r.stderr              # EB:947-950
    r = run(cli, "--num-layers", "3", "--heads", "4,1", "--outdims", "8,8,8")
    assert r.returncode == 1 and "Error: --heads must have 3 values." in r.stderr               # EB:969-972
    r = run(cli, "--num-layers", "2", "--heads", "4,1", "--outdims", "8")
    assert r.returncode == 1 and "must have 2 values." in r.stderr                              # EB:981-984
    r = run(cli, "--heads", "4,1", "--outdims", "8,8", "--optimizer", "rmsprop")
    assert r.returncode == 1 and "Invalid optimizer choice. Use 'sgd' or 'adam'" in r.stderr    # EB:993-996
    r = run(cli, "--heads", "4,1", "--outdims", "8,8", "--optimizer", "adam", "--beta2", "1.0")
    assert r.returncode == 1 and "beta1 and beta2 must be in (0,1)" in r.stderr                 # EB:1011-1015
    r = run(cli, "--heads", "4,1")
    assert r.returncode == 1 and "--heads and --outdims must be given" in r.stderr


def test_config_block_and_missing_dataset(cli, tmp_path):
    r = run(cli, "--bogus-flag", "--num-layers", "3", "--heads", "4,1,1", "--outdims", "64,32,16", "--epochs", "7",
            "--optimizer", "sgd", "--beta1", "0.5", "--lr", "0.01", "--clip", "--dataset", "citeseer", "--data-root",
            str(tmp_path))
    assert "Warning: beta1/beta2 specified but ignored for SGD optimizer." in r.stderr          # EB:1016-1019
    expect = ("Configuration:\n  Number of layers: 3\n  Epochs: 7\n  Attention heads: [4, 1, 1]\n"
              "  Output dimensions: [64, 32, 16]\n  Gradient clipping: true\n  Optimizer: sgd\n"
              "  Learning rate: 0.01\n\nUsing dataset: citeseer\nDataset path: %s/citeseer/\n" % tmp_path)
    assert r.stdout.startswith(expect), r.stdout                                                 # EB:1024-1040, 1076-1077
    assert r.returncode == 1 and "Invalid row_ptr length" in r.stderr                            # EB:1084-1087
    env = dict(os.environ, DATA_ROOT=str(tmp_path / "elsewhere"))
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", env=env)
    assert "Using dataset: pubmed\nDataset path: %s/elsewhere/pubmed/\n" % tmp_path in r.stdout  # EB:1050, 1064-1073


def test_inconsistent_feature_line(cli, tmp_path):
    d = tmp_path / "bad"
    d.mkdir()
    (d / "features.txt").write_text("1 2 3\n4 5\n")
    for f in ("row_ptr.txt", "col_idx.txt", "labels.txt"):
        (d / f).write_text("0\n")
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "bad", "--data-root", str(tmp_path))
    assert r.returncode == 1 and "Inconsistent input_dim on line 1" in r.stderr                  # EB:42-45


@pytest.mark.gpu
@pytest.mark.parametrize("name,gemm,tol", [("sample_adam", "fp32", 5e-4), ("three_layer_adam", "fp32", 5e-4),
                                           ("sample_adam", "tf32", 1e-2)])
def test_cli_reproduces_reference_loss_curve(cli, tmp_path, name, gemm, tol):
    sys.path.insert(0, PKG)
    import datasets
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_edge_%s.npz" % name))
    heads, outdims = g["heads"].tolist(), g["outdims"].tolist()
    datasets.write_txt(str(tmp_path / "data" / name), dict(X=g["X"], row_ptr=g["row_ptr"], col_idx=g["col_idx"],
                                                          labels=g["labels"]))
    w = tmp_path / "w"
    w.mkdir()
    np.concatenate([g["W_%d" % l].ravel() for l in range(len(heads))]).tofile(str(w / "W.bin"))
    np.concatenate([g["a_%d" % l].ravel() for l in range(len(heads))]).tofile(str(w / "a.bin"))
    g["Wo"].astype(np.float32).tofile(str(w / "Wo.bin"))
    curve = g["loss_curve"]
    args = ["--num-layers", str(len(heads)), "--heads", ",".join(map(str, heads)), "--outdims",
            ",".join(map(str, outdims)), "--epochs", str(len(curve)), "--optimizer", str(g["optimizer"]), "--lr",
            str(float(g["lr"])), "--dataset", name, "--data-root", str(tmp_path / "data"), "--load-weights", str(w),
            "--gemm", gemm] + (["--clip"] if bool(g["clip"]) else [])
    r = run(cli, *args)
    assert r.returncode == 0, r.stderr
    assert "Max degree = %d\n" % int(g["max_degree"]) in r.stdout and "Number of classes = %d\n" % int(g["num_classes"]) in r.stdout
    assert "Graph loaded: %d nodes, %d edges, input_feature_vector_dim = %d\n" % (len(g["labels"]), len(g["col_idx"]), g["X"].shape[1]) in r.stdout
    got = re.findall(r"\nEpoch (\d+)\n\nAvg Loss: ([0-9.]+), Accuracy: ([0-9.]+)%\n total time: [0-9.e+-]+ ms\n", r.stdout)
    assert len(got) == len(curve), r.stdout[-800:]
    for (ep, loss, acc), (rl, ra) in zip(got, curve):
        assert abs(float(loss) - rl) < tol * max(1.0, rl), (ep, loss, rl)
        assert abs(float(acc) - ra) <= 100.0 / len(g["labels"]) + 0.011, (ep, acc, ra)
