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


def test_binary_dataset_cache(cli, tmp_path):
    """--load-only: the first run parses the text files and writes <dataset>/.gatx_cache.bin, the second run reads
    it; both print the same dataset lines; editing a text file invalidates the cache (size / mtime key)."""
    sys.path.insert(0, PKG)
    import datasets
    ds = datasets.make_dataset("sample")
    d = tmp_path / "sample"
    datasets.write_txt(str(d), ds)
    args = ["--heads", "8,1", "--outdims", "8,8", "--dataset", "sample", "--data-root", str(tmp_path), "--load-only"]
    r1 = run(cli, *args)
    assert r1.returncode == 0 and "[loader] text files" in r1.stderr and (d / ".gatx_cache.bin").exists()
    r2 = run(cli, *args)
    assert r2.returncode == 0 and "[loader] binary cache" in r2.stderr and r2.stdout == r1.stdout
    assert "Max degree = %d\n" % int(np.diff(ds["row_ptr"]).max()) in r2.stdout
    assert "Graph loaded: 64 nodes, 512 edges, input_feature_vector_dim = 16\n" in r2.stdout
    lab = ds["labels"].copy()
    lab[0] = 9  # a new class: file size and class count change
    np.savetxt(str(d / "labels.txt"), np.concatenate([lab, []]).astype(int), fmt="%d")
    open(str(d / "labels.txt"), "a").write("\n")
    r3 = run(cli, *args)
    assert "[loader] text files" in r3.stderr and "Number of classes = 10\n" in r3.stdout
    r4 = run(cli, *(args + ["--no-cache"]))
    assert "[loader] text files" in r4.stderr and r4.stdout == r3.stdout


def _read_cache(path):
    """<dataset>/.gatx_cache.bin: header (magic, N, I, E, 4 sizes, 4 mtimes) then X, row_ptr, col_idx, labels."""
    raw = open(path, "rb").read()
    assert raw[:7] == b"GATXDS1"
    N, I, E = np.frombuffer(raw, np.int64, 3, 8)
    off = 8 + 8 * 3 + 8 * 8
    X = np.frombuffer(raw, np.float32, N * I, off).reshape(N, I)
    off += 4 * N * I
    rp = np.frombuffer(raw, np.int32, N + 1, off)
    off += 4 * (N + 1)
    ci = np.frombuffer(raw, np.int32, E, off)
    off += 4 * E
    return X, rp, ci, np.frombuffer(raw, np.int32, N, off)


@pytest.mark.parametrize("threads", ["1", "5"])
def test_parallel_text_loader_is_bit_exact(cli, tmp_path, threads):
    """The text files are cut into one segment per host thread (lines for features.txt, tokens for the integer
    files); whatever the thread count, the parsed arrays are the bits numpy reads from the same files, incl. tokens
    only strtof understands ("+1.5", "inf", "1e-3", "0x1p-2") and a last line without a newline."""
    sys.path.insert(0, PKG)
    import datasets
    rng = np.random.default_rng(5)
    N, I = 30000, 40   # ~13 MB of text: several 4 MB segments
    X = (rng.standard_normal((N, I)) * 10.0 ** rng.integers(-20, 20, (N, I))).astype(np.float32)
    row_ptr, col_idx = datasets.make_graph(N, 200000, "rmat", 3)
    labels = rng.integers(0, 7, N).astype(np.int32)
    d = tmp_path / "big"
    datasets.write_txt(str(d), dict(X=X, row_ptr=row_ptr, col_idx=col_idx, labels=labels))
    lines = open(d / "features.txt").read().split("\n")
    assert lines[-1] == ""
    special = ["+1.5", "inf", "-inf", "1e-3", "0x1p-2", "1E+5", ".5", "5."]
    toks = lines[12345].split(" ")
    toks[:len(special)] = special
    lines[12345] = "  ".join(toks) + " \t"          # extra blanks and trailing whitespace
    X[12345, :len(special)] = [1.5, np.inf, -np.inf, 1e-3, 0.25, 1e5, 0.5, 5.0]
    open(d / "features.txt", "w").write("\n".join(lines[:-1]))  # no newline after the last line
    env = dict(os.environ, GATX_LOADER_THREADS=threads)
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "big", "--data-root", str(tmp_path), "--load-only",
            env=env)
    assert r.returncode == 0, r.stderr
    assert "Graph loaded: %d nodes, %d edges, input_feature_vector_dim = %d\n" % (N, len(col_idx), I) in r.stdout
    Xc, rp, ci, lab = _read_cache(d / ".gatx_cache.bin")
    assert np.array_equal(Xc.view(np.uint32), X.view(np.uint32))     # bit-exact floats
    assert np.array_equal(rp, row_ptr) and np.array_equal(ci, col_idx) and np.array_equal(lab, labels)
    # a short line deep inside a later segment is reported with its line number (EB:42-45)
    lines[20001] = " ".join(lines[20001].split()[:-1])
    open(d / "features.txt", "w").write("\n".join(lines))
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "big", "--data-root", str(tmp_path), "--load-only",
            "--no-cache", env=env)
    assert r.returncode == 1 and "Inconsistent input_dim on line 20001\n" in r.stderr
    # `file >> int` stops at the first token that is not an integer (EB:53-64): col_idx is short and the run refuses it
    open(d / "features.txt", "w").write("\n".join(lines[:20001] + [lines[0]] + lines[20002:]))
    ctoks = open(d / "col_idx.txt").read().split()
    ctoks[150000] = "x7"
    open(d / "col_idx.txt", "w").write("\n".join(ctoks) + "\n")
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "big", "--data-root", str(tmp_path), "--load-only",
            "--no-cache", env=env)
    assert r.returncode == 1 and "Invalid col_idx length" in r.stderr


@pytest.mark.gpu
def test_checkpoint_resume_is_bit_exact(cli, tmp_path):
    """6 epochs straight == 3 epochs + --save-checkpoint, then --resume for 3 more (Adam moments and t restored)."""
    sys.path.insert(0, PKG)
    import datasets
    ds = datasets.make_dataset("sample")
    datasets.write_txt(str(tmp_path / "sample"), ds)
    base = ["--heads", "8,1", "--outdims", "8,8", "--optimizer", "adam", "--lr", "0.01", "--clip", "--dataset", "sample",
            "--data-root", str(tmp_path), "--seed", "5"]
    for name in ("a", "b"):
        (tmp_path / name).mkdir()
    r = run(cli, *(base + ["--epochs", "6", "--dump-weights", str(tmp_path / "a")]))
    assert r.returncode == 0, r.stderr
    r = run(cli, *(base + ["--epochs", "3", "--save-checkpoint", str(tmp_path / "ck.bin")]))
    assert r.returncode == 0, r.stderr
    r = run(cli, *(base + ["--epochs", "3", "--resume", str(tmp_path / "ck.bin"), "--dump-weights", str(tmp_path / "b")]))
    assert r.returncode == 0 and "\nEpoch 4\n" in r.stdout and "\nEpoch 6\n" in r.stdout, r.stderr
    for f in ("W.bin", "a.bin", "Wo.bin"):
        assert (tmp_path / "a" / f).read_bytes() == (tmp_path / "b" / f).read_bytes(), f


@pytest.mark.gpu
def test_checkpoint_of_another_model_is_refused(cli, tmp_path):
    """4 heads x 16 and 8 heads x 8 have the same number of parameters: the checkpoint header (magic, version, layers,
    heads, outdims, in_dim, classes, bias, optimizer) must refuse the file instead of reinterpreting the floats."""
    sys.path.insert(0, PKG)
    import datasets
    datasets.write_txt(str(tmp_path / "sample"), datasets.make_dataset("sample"))
    common = ["--optimizer", "adam", "--lr", "0.01", "--dataset", "sample", "--data-root", str(tmp_path), "--seed", "5",
              "--epochs", "2"]
    ck = str(tmp_path / "ck.bin")
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", *common, "--save-checkpoint", ck)
    assert r.returncode == 0, r.stderr
    r = run(cli, "--heads", "4,1", "--outdims", "16,8", *common, "--resume", ck)
    assert r.returncode == 1 and "written for another model" in r.stderr
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", *[("sgd" if x == "adam" else x) for x in common], "--resume", ck)
    assert r.returncode == 1 and "another optimizer" in r.stderr
    open(tmp_path / "junk.bin", "wb").write(b"\0" * 4096)
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", *common, "--resume", str(tmp_path / "junk.bin"))
    assert r.returncode == 1 and "not a gatx checkpoint" in r.stderr
    # --bias: the biases travel with --dump-weights / --load-weights (b.bin)
    (tmp_path / "w").mkdir()
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", *common, "--bias", "--dump-weights", str(tmp_path / "w"))
    assert r.returncode == 0 and (tmp_path / "w" / "b.bin").stat().st_size == 4 * (64 + 8)
    b = np.fromfile(tmp_path / "w" / "b.bin", np.float32)
    assert np.abs(b).max() > 0  # trained for two epochs
    (tmp_path / "w2").mkdir()
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", *common[:-1], "0", "--bias", "--load-weights", str(tmp_path / "w"),
            "--dump-weights", str(tmp_path / "w2"))
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "w2" / "b.bin").read_bytes() == (tmp_path / "w" / "b.bin").read_bytes()


@pytest.mark.gpu
def test_extension_flags(cli, tmp_path):
    """--attn-slope / --act-slope / --dropout are opt-in: absent (or at the reference values) the loss curve is the
    default one bit for bit; present they change it, reproducibly for a fixed --seed."""
    sys.path.insert(0, PKG)
    import datasets
    datasets.write_txt(str(tmp_path / "sample"), datasets.make_dataset("sample"))
    base = ["--heads", "8,1", "--outdims", "8,8", "--optimizer", "adam", "--lr", "0.01", "--dataset", "sample",
            "--data-root", str(tmp_path), "--seed", "5", "--epochs", "5"]

    def losses(*extra):
        r = run(cli, *(base + list(extra)))
        assert r.returncode == 0, r.stderr
        return re.findall(r"Avg Loss: ([0-9.]+)", r.stdout), r.stdout

    default, out = losses()
    assert len(default) == 5 and "Dropout" not in out
    assert losses("--attn-slope", "0.01", "--act-slope", "0.01", "--dropout", "0")[0] == default
    slope, _ = losses("--attn-slope", "0.2")
    assert slope != default and slope[0] != default[0]  # the very first forward already differs
    drop, out = losses("--dropout", "0.4")
    assert "Dropout: 0.4" in out and drop != default
    assert losses("--dropout", "0.4")[0] == drop
    bias, _ = losses("--bias")
    assert bias[0] == default[0] and bias[1:] != default[1:]  # biases start at zero, then train
    r = run(cli, *(base + ["--dropout", "1.5"]))
    assert r.returncode == 1 and "gatx_set_dropout" in r.stderr


def test_invalid_label_and_gpu_count(cli, tmp_path):
    """A negative label is refused while loading (the reference would index its class arrays out of bounds, EB:524);
    --gpus N with fewer usable devices ends with an error instead of leaving rank threads waiting in NCCL."""
    sys.path.insert(0, PKG)
    import datasets
    ds = datasets.make_dataset("sample")
    ds["labels"] = ds["labels"].copy()
    ds["labels"][10] = -2
    datasets.write_txt(str(tmp_path / "neg"), ds)
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "neg", "--data-root", str(tmp_path), "--load-only")
    assert r.returncode == 1 and "Invalid label on line 11" in r.stderr
    datasets.write_txt(str(tmp_path / "ok"), datasets.make_dataset("sample"))
    r = run(cli, "--heads", "8,1", "--outdims", "8,8", "--dataset", "ok", "--data-root", str(tmp_path), "--gpus", "64",
            "--epochs", "1")
    assert r.returncode == 1 and "usable CUDA device" in r.stderr


def test_split_file_errors(cli, tmp_path):
    """--split: <dataset>/split.txt must hold one token in {0, 1, 2} per node (checked before any GPU work)."""
    sys.path.insert(0, PKG)
    import datasets
    ds = datasets.make_dataset("sample")
    d = tmp_path / "sample"
    datasets.write_txt(str(d), ds)
    args = ["--heads", "8,1", "--outdims", "8,8", "--dataset", "sample", "--data-root", str(tmp_path), "--split", "--no-cache"]
    (d / "split.txt").write_text("0\n1\n2\n")
    r = run(cli, *args)
    assert r.returncode == 1 and "Invalid split length" in r.stderr
    (d / "split.txt").write_text("\n".join(["0"] * 10 + ["3"] + ["1"] * 53) + "\n")
    r = run(cli, *args)
    assert r.returncode == 1 and "Invalid split value on line 11" in r.stderr


@pytest.mark.gpu
def test_split_training_and_eval_only_match_oracle(cli, tmp_path):
    """--split trains on the train nodes and reports validation / test metrics; --eval-only re-scores the dumped
    weights.  The printed numbers are checked against the oracle's masked variant with the same injected weights."""
    sys.path.insert(0, PKG)
    import datasets
    import orc
    from helpers import make_oracle, make_problem
    p = make_problem(64, 512, 16, 4, (8, 1), (8, 8), "uniform", seed=11)
    d = tmp_path / "data" / "s"
    datasets.write_txt(str(d), dict(X=p["X"], row_ptr=p["row_ptr"], col_idx=p["col_idx"], labels=p["labels"]))
    part = np.random.default_rng(4).integers(0, 3, 64)
    np.savetxt(str(d / "split.txt"), part, fmt="%d")
    w = tmp_path / "w"
    w.mkdir()
    np.concatenate([x.ravel() for x in p["Ws"]]).astype(np.float32).tofile(str(w / "W.bin"))
    np.concatenate([x.ravel() for x in p["As"]]).astype(np.float32).tofile(str(w / "a.bin"))
    p["Wo"].astype(np.float32).tofile(str(w / "Wo.bin"))
    out = tmp_path / "out"
    out.mkdir()
    base = ["--heads", "8,1", "--outdims", "8,8", "--optimizer", "adam", "--lr", "0.01", "--dataset", "s", "--data-root",
            str(tmp_path / "data"), "--gemm", "fp32", "--split"]
    r = run(cli, *(base + ["--epochs", "5", "--load-weights", str(w), "--dump-weights", str(out)]))
    assert r.returncode == 0, r.stderr
    masks = [(part == k).astype(np.uint8) for k in range(3)]
    ref = make_oracle(orc, p, optimizer="adam", lr=0.01)
    got = re.findall(r"\nAvg Loss: ([0-9.]+), Accuracy: ([0-9.]+)%\n\nVal Loss: ([0-9.]+), Val Accuracy: ([0-9.]+)%\n", r.stdout)
    assert len(got) == 5, r.stdout[-600:]

    def score(mask):
        ref.set_mask(mask)
        ref.forward()
        o = ref.loss()
        return o["avg"], 100.0 * o["acc"]

    for t, (tl, ta, vl, va) in enumerate(got, 1):
        ref.set_mask(masks[0])
        ol, oa = ref.epoch(t)
        assert abs(float(tl) - ol) < 5e-4 * max(1.0, ol) and abs(float(ta) - 100.0 * oa) < 0.011, (t, tl, ol)
        rvl, rva = score(masks[1])
        assert abs(float(vl) - rvl) < 5e-4 * max(1.0, rvl) and abs(float(va) - rva) < 0.011, (t, vl, rvl)
    m = re.search(r"\nTest Loss: ([0-9.]+), Test Accuracy: ([0-9.]+)%\n", r.stdout)
    rtl, rta = score(masks[2])
    assert m and abs(float(m.group(1)) - rtl) < 5e-4 * max(1.0, rtl) and abs(float(m.group(2)) - rta) < 0.011
    # evaluation only, from the dumped weights: the three splits, then all nodes without --split
    r2 = run(cli, *(base + ["--eval-only", "--load-weights", str(out)]))
    assert r2.returncode == 0 and "\nEpoch " not in r2.stdout, r2.stderr
    for k, name in enumerate(("Train", "Val", "Test")):
        m = re.search(r"\n%s Loss: ([0-9.]+), %s Accuracy: ([0-9.]+)%%\n" % (name, name), r2.stdout)
        el, ea = score(masks[k])
        assert m and abs(float(m.group(1)) - el) < 5e-4 * max(1.0, el) and abs(float(m.group(2)) - ea) < 0.011, name
    r3 = run(cli, *([a for a in base if a != "--split"] + ["--eval-only", "--load-weights", str(out)]))
    el, ea = score(None)
    m = re.search(r"\nAvg Loss: ([0-9.]+), Accuracy: ([0-9.]+)%\n", r3.stdout)
    assert r3.returncode == 0 and m and abs(float(m.group(1)) - el) < 5e-4 * max(1.0, el)


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
