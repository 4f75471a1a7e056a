"""Build libtsim.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m text_similarity_b200.build [--force] [--verbose] [--experiment]

Every .cu is compiled to an object in parallel (one nvcc per file), then linked.  The .so is git-ignored
but travels to the GPU box with the repo snapshot.  ``--experiment`` builds libtsim_exp.so with
-DTSIM_EXPERIMENT (environment knobs and in-kernel diagnosis switches compiled in) for scripts/ab_*.py;
the package itself only ever loads the release library.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtsim.so")
LIB_EXPERIMENT = os.path.join(HERE, "libtsim_exp.so")
OBJ_DIR = os.path.join(HERE, "build")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def _newer(path: str, than: float) -> bool:
    return os.path.getmtime(path) > than


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(_newer(d, t) for d in sources() + _deps())


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libtsim.so cannot be built on this machine")
    return nvcc


def build(force: bool = False, verbose: bool = False, experiment: bool = False) -> str:
    lib = LIB_EXPERIMENT if experiment else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = find_nvcc()
    flags = NVCC_FLAGS + (["-DTSIM_EXPERIMENT"] if experiment else []) + (["-Xptxas", "-v"] if verbose else [])
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(d) for d in _deps())
    tag = "exp" if experiment else "rel"

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, f"{os.path.splitext(os.path.basename(src))[0]}.{tag}.o")
        if (not force and os.path.exists(obj) and not _newer(src, os.path.getmtime(obj))
                and os.path.getmtime(obj) > hdr_time):
            return obj, None
        res = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    failed = False
    for obj, res in results:
        if res is None:
            continue
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        failed = failed or res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libtsim.so")
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib]
                         + [obj for obj, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libtsim.so")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, experiment="--experiment" in sys.argv))
