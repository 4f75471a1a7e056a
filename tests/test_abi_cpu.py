"""CPU-side checks: libtsim.so builds/loads and exports every symbol include/tsim.h declares;
host-only entry points behave (no kernel is launched here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from text_similarity_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from text_similarity_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tsim.h")).read()
    declared = set(re.findall(r"\b(tsim_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_workspace_planning(lib):
    from text_similarity_b200 import _lib
    assert lib.tsim_version() == _lib.ABI_VERSION == 2
    # BASELINE configs: workspace stays a small fraction of the 180 GB of HBM
    for (Q, N, D, k) in [(100, 10_000, 384, 10), (1024, 1_000_000, 768, 10), (4096, 1_250_000, 768, 100),
                         (32, 12_500_000, 384, 10), (1, 10_000_000, 768, 10)]:
        nb = lib.tsim_search_workspace_bytes(Q, N, D, k, _lib.BF16, _lib.BF16, _lib.MODE_AUTO)
        assert 0 < nb < 8e9, (Q, N, D, k, nb)
    assert lib.tsim_search_workspace_bytes(8, 100, 30, 5, _lib.F32, _lib.F32, _lib.MODE_TENSOR) == 0
    assert b"TSIM_MODE_TENSOR" in lib.tsim_last_error()
    assert lib.tsim_search_workspace_bytes(8, 100, 32, 0, _lib.F32, _lib.F32, _lib.MODE_AUTO) == 0
    assert lib.tsim_pool_workspace_bytes(16, 256, 384) > 0


def test_planner_sweep_over_legal_shapes_never_wraps_or_crashes(lib):
    """Host-side planning (no GPU needed: 148 SMs assumed) over random legal shapes up to the ABI's limits: the
    workspace size is positive, holds at least the float64 fallback lists (Q * k * 12 bytes: catches 32-bit
    wrap-around in the offset arithmetic); illegal shapes are errors with a message, never a crash."""
    import numpy as np
    from text_similarity_b200 import _lib
    rng = np.random.default_rng(7)
    dts = [_lib.BF16, _lib.E4M3, _lib.F32, _lib.F16]
    for _ in range(3000):
        Q = int(rng.choice([1, 2, 31, 32, 33, 128, 129, 4096, 100_000, 2 ** 31 - 1]))
        N = int(rng.choice([0, 1, 255, 256, 257, 303_103, 303_104, 10_000_000, 100_000_000, 4_000_000_000]))
        D = int(rng.choice([8, 16, 24, 64, 384, 768, 1000, 4096, 65_536]))
        k = int(rng.choice([1, 10, 11, 24, 25, 52, 53, 100, 101, 1024]))
        qd, cd = int(rng.choice(dts)), int(rng.choice(dts))
        mode = int(rng.choice([_lib.MODE_AUTO, _lib.MODE_EXACT]))
        nb = lib.tsim_search_workspace_bytes(Q, N, D, k, qd, cd, mode)
        assert nb >= Q * k * 12, (Q, N, D, k, qd, cd, mode, nb, lib.tsim_last_error())
        if D % 8 == 0 and N > 0:
            for sdt in (_lib.BF16,):
                ns = lib.tsim_search_shadow_workspace_bytes(Q, N, D, k, sdt)
                assert ns >= Q * k * 12, (Q, N, D, k, ns, lib.tsim_last_error())
    for bad in [(-1, 10, 8, 1), (1, -1, 8, 1), (1, 10, 0, 1), (1, 10, 8, 0), (1, 10, 8, 1025), (1, 2 ** 32, 8, 1), (2 ** 31, 10, 8, 1)]:
        assert lib.tsim_search_workspace_bytes(*bad, _lib.BF16, _lib.BF16, _lib.MODE_AUTO) == 0
        assert len(lib.tsim_last_error()) > 0


def test_argument_errors_do_not_touch_the_gpu(lib):
    from text_similarity_b200 import _lib
    rc = lib.tsim_merge_topk(None, None, 4, 100, 100, 10, None, None, None, None)
    assert rc == _lib.ERR_INVALID_ARG and b"4096" in lib.tsim_last_error()
    rc = lib.tsim_row_inv_norm(None, _lib.F32, 10, 0, 0, None, None)
    assert rc == _lib.ERR_INVALID_ARG
    rc = lib.tsim_pool_norm(None, _lib.E4M3, None, _lib.I64, 2, 3, 4, 12, 4, 3, None, _lib.F32, 4, None, None, 1, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARG


def test_ops_refuse_cpu_tensors():
    import torch
    from text_similarity_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.search_topk(torch.randn(2, 8), torch.randn(5, 8), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pool_norm(torch.randn(2, 3, 8), torch.ones(2, 3, dtype=torch.int64))


def test_plan_handle_host_side(lib):
    """tsim_plan_create / _workspace_bytes / _destroy are host-only: same workspace size as the plain query
    function, NULL + message on an unsupported shape, and no environment reads / descriptor encodes."""
    from text_similarity_b200 import _lib
    before = _lib.counters()
    h = lib.tsim_plan_create(4096, 1_250_000, 768, 100, _lib.BF16, _lib.BF16, _lib.MODE_AUTO, -1)
    assert h
    assert lib.tsim_plan_workspace_bytes(h) == lib.tsim_search_workspace_bytes(
        4096, 1_250_000, 768, 100, _lib.BF16, _lib.BF16, _lib.MODE_AUTO)
    lib.tsim_plan_destroy(h)
    hs = lib.tsim_plan_create(100, 10_000, 384, 10, _lib.F32, _lib.F32, _lib.MODE_AUTO, _lib.BF16)
    assert hs
    assert lib.tsim_plan_workspace_bytes(hs) == lib.tsim_search_shadow_workspace_bytes(100, 10_000, 384, 10, _lib.BF16)
    lib.tsim_plan_destroy(hs)
    assert not lib.tsim_plan_create(8, 100, 30, 5, _lib.F32, _lib.F32, _lib.MODE_TENSOR, -1)
    assert b"TSIM_MODE_TENSOR" in lib.tsim_last_error()
    assert not lib.tsim_plan_create(8, 100, 32, 5, _lib.F32, _lib.F32, _lib.MODE_AUTO, _lib.F16)   # shadow must be bf16
    assert lib.tsim_plan_workspace_bytes(None) == 0
    lib.tsim_plan_destroy(None)
    after = _lib.counters()
    assert after["env_reads"] == before["env_reads"] == 0
    assert after["map_encodes"] == before["map_encodes"]
    assert after["plans"] > before["plans"]
    rc = lib.tsim_plan_search(None, None, 0, None, 0, None, None, 0, None, 0, 0, -1, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARG


def test_release_library_has_no_experiment_knobs(lib, monkeypatch):
    """The library the package loads reads no environment variable: the knob names are not even in the binary
    (they exist only in libtsim_exp.so, built with -DTSIM_EXPERIMENT for scripts/ab_*.py), and planning is the
    same whatever TSIM_* variables say."""
    from text_similarity_b200 import _lib
    assert lib.tsim_build_flags() == 0
    blob = open(_lib.LIB_PATH, "rb").read()
    for knob in (b"TSIM_DEBUG", b"TSIM_NO_PAIR", b"TSIM_NO_BOOT", b"TSIM_HOT_SCALED", b"TSIM_ROLES_LOW",
                 b"TSIM_STATIC_UNITS", b"TSIM_POOL_DEBUG", b"TSIM_CHUNK_ROWS", b"TSIM_NO_RETRY", b"TSIM_NO_LADDER"):
        assert knob not in blob, knob
    args = (4096, 10_000_000, 768, 10, _lib.BF16, _lib.BF16, _lib.MODE_AUTO)
    base = lib.tsim_search_workspace_bytes(*args)
    for name, val in (("TSIM_DEBUG", "2"), ("TSIM_NO_PAIR", "1"), ("TSIM_NO_BOOT", "1"), ("TSIM_CHUNK_ROWS", "65536")):
        monkeypatch.setenv(name, val)
    assert lib.tsim_search_workspace_bytes(*args) == base
    assert _lib.counters()["env_reads"] == 0
