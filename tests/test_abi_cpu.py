"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/gatx.h declares,
refuses to create a context without a CUDA device (no CPU fallback), and its host-side
partitioning is bit-exact against the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import build as gatx_build
    import gatx
    gatx_build.build()
    return gatx.load()


def test_exports_every_declared_symbol(lib):
    import gatx
    hdr = open(os.path.join(ROOT, "include", "gatx.h")).read()
    declared = set(re.findall(r"\b(gatx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gatx_ctx", "gatx_config", "gatx_status"}
    assert declared == set(gatx.EXPORTS), declared ^ set(gatx.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gatx_version()


def test_no_cpu_fallback(lib):
    import gatx
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gatx.GatxError):
        gatx.Engine([8, 1], [8, 8])


@pytest.mark.parametrize("R", [1, 2, 4, 8])
def test_partition_rows_bit_exact(lib, orc, R):
    import gatx
    row_ptr, _ = datasets.make_graph(3000, 40000, "rmat", 11)
    assert np.array_equal(gatx.partition_rows(row_ptr, R), orc.partition_rows(row_ptr, R))


@pytest.mark.parametrize("R", [1, 2, 3, 8])
@pytest.mark.parametrize("K", [1, 3, 4])
def test_row_blocks_bit_exact(lib, orc, R, K):
    """Block table of the pipelined multi-GPU exchange: every rank derives the same table from the global row_ptr, so it
    must be bit-exact (against an independent numpy formulation), cover every own row exactly once and balance EDGES."""
    import gatx
    for N, E, kind in ((3000, 40000, "rmat"), (97, 97, "uniform"), (500, 9000, "uniform")):
        row_ptr, _ = datasets.make_graph(N, E, kind, 11)
        blk = gatx.row_blocks(row_ptr, R, K)
        assert np.array_equal(blk, orc.row_blocks(row_ptr, R, K))
        bounds = orc.partition_rows(row_ptr, R)
        assert np.array_equal(blk[:, 0], bounds[:-1]) and np.array_equal(blk[:, -1], bounds[1:])
        assert (np.diff(blk, axis=1) >= 0).all()
        if kind == "uniform" and N == 500:  # a flat degree distribution: blocks within two rows' worth of the ideal
            edges = np.diff(np.asarray(row_ptr, np.int64)[blk], axis=1)
            ideal = edges.sum(1, keepdims=True) / K
            assert (np.abs(edges - ideal) <= 2 * np.diff(row_ptr).max() + 1).all()


def test_sass_is_sm100a_only(lib):
    import subprocess
    import gatx
    out = subprocess.run(["cuobjdump", "-lelf", gatx.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
