"""Builds libgatx.so (the C-ABI CUDA library) in-tree for sm_100a.  `python build.py [--force]`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# GATX_VARIANT=<name> GATX_EXTRA_FLAGS="-D..." builds an A/B variant libgatx_<name>.so next to the product library
VARIANT = os.environ.get("GATX_VARIANT", "")
OUT = os.path.join(HERE, "libgatx_%s.so" % VARIANT if VARIANT else "libgatx.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"] + os.environ.get("GATX_EXTRA_FLAGS", "").split()
SOURCES = ["gatx_api.cu", "graph_prep.cu", "gemm_simt.cu", "gemm_tc.cu", "edge_kernels.cu", "edge_stream.cu", "edge_generic.cu", "head_loss.cu", "optim.cu", "halo_p2p.cu"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build_" + VARIANT if VARIANT else "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "gatx.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd[-3:]) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart", "-ldl", "-ccbin", "/usr/bin/g++"]
        subprocess.check_call(cmd)
    return OUT


def build_cli():
    """train_gatx: the reference-compatible command line (host C++ above the C ABI)."""
    src = os.path.join(CSRC, "train_main.cpp")
    out = os.path.join(HERE, "train_gatx")
    if _stale(out, [src, OUT, os.path.join(HERE, "..", "include", "gatx.h")]):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", src, "-o", out, "-L" + HERE, "-lgatx",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + "/usr/local/cuda/lib64", "-lpthread"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli())
