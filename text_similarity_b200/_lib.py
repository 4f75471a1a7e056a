"""ctypes binding of libtsim.so -- the C ABI declared in include/tsim.h.

There is deliberately no fallback: if the library is missing or a call fails, an exception
is raised.  The product path never computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtsim.so")
EXPERIMENT_LIB_PATH = os.path.join(HERE, "libtsim_exp.so")   # -DTSIM_EXPERIMENT build, scripts/ab_*.py only
ABI_VERSION = 2

# element-type codes (include/tsim.h)
F32, F16, BF16, E4M3 = 0, 1, 2, 3
I64, I32, U8 = 10, 11, 12
MODE_AUTO, MODE_EXACT, MODE_TENSOR = 0, 1, 2
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_MISALIGNED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4, -5

# every symbol include/tsim.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tsim_version": (c_int, []),
    "tsim_last_error": (c_char_p, []),
    "tsim_pool_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "tsim_pool_norm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64,
                               c_int64, c_int64, c_int64, c_void_p, c_int, c_int64, c_void_p,
                               c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "tsim_row_inv_norm": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "tsim_search_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "tsim_search_topk": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p,
                                 c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tsim_search_shadow_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "tsim_search_topk_shadow": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64,
                                        c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p,
                                        c_int64, c_int64, c_int64, c_int, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tsim_plan_create": (c_void_p, [c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int]),
    "tsim_plan_create_split_shadow": (c_void_p, [c_int64, c_int64, c_int64, c_int, c_int, c_int]),
    "tsim_plan_destroy": (None, [c_void_p]),
    "tsim_plan_workspace_bytes": (c_size_t, [c_void_p]),
    "tsim_plan_search": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tsim_debug_counters": (None, [POINTER(ctypes.c_uint64)]),
    "tsim_build_flags": (c_int, []),
    "tsim_debug_eps": (c_float, [c_int64, c_int, c_int]),
    "tsim_debug_tensor_pass": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64,
                                       c_int64, c_void_p, POINTER(c_int64), c_void_p, c_void_p]),
    "tsim_set_timing_events": (c_int, [c_void_p, c_void_p]),
    "tsim_launch_count": (ctypes.c_uint64, []),
    "tsim_merge_topk_strided": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "tsim_merge_topk": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class TsimError(RuntimeError):
    """A libtsim call failed (CUDA error or workspace problem)."""


def use_experiment_build() -> None:
    """Measurement scripts only: bind the -DTSIM_EXPERIMENT flavour (environment knobs compiled in) instead
    of the release library.  Must be called before the first load()."""
    global LIB_PATH
    if _lib is not None:
        raise TsimError("use_experiment_build() must precede the first load()")
    LIB_PATH = EXPERIMENT_LIB_PATH


def counters() -> dict:
    out = (ctypes.c_uint64 * 4)()
    load().tsim_debug_counters(out)
    return {"launches": out[0], "map_encodes": out[1], "env_reads": out[2], "plans": out[3]}


def load() -> ctypes.CDLL:
    """Load libtsim.so (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TsimError(
            f"{LIB_PATH} is missing: build it with `python -m text_similarity_b200.build` "
            "(there is no CPU fallback for the search path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tsim_version() != ABI_VERSION:
        raise TsimError(f"libtsim ABI version {lib.tsim_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == OK:
        return
    msg = load().tsim_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg} (code {rc})"
    if rc in (ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_MISALIGNED):
        raise ValueError(text)
    raise TsimError(text)
