"""A short slice of the randomized sweeps under scripts/ (fuzz_parity.py, fuzz_pool.py, fuzz_merge.py) inside the GPU
suite: seeded, a fixed number of cases (the same ones every run), run as the scripts themselves so that a replay
(`--cases`) of anything they report uses the same code.  The long runs are recorded in profiles/r02c_fuzz.txt."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, n_cases, *args):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    args = ("--seconds", "400", "--max-cases", str(n_cases)) + args
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-15:])
    assert res.returncode == 0, f"{script} {' '.join(args)} reported mismatches:\n{tail}"
    last = res.stdout.strip().splitlines()[-1]
    assert " 0 failing" in last, last
    cases = int(last.split(":")[1].split("cases")[0])
    assert cases == n_cases, f"{cases} of {n_cases} cases ran: {last}"


def test_search_whatever_the_planner_picks_equals_the_float64_scan():
    _run("fuzz_parity.py", 600, "--seed", "101")


def test_search_big_shapes_equal_the_float64_scan():
    _run("fuzz_parity.py", 30, "--seed", "102", "--big")


def test_pool_equals_the_float64_formula():
    _run("fuzz_pool.py", 2000, "--seed", "101")


def test_merge_equals_a_stable_sort():
    _run("fuzz_merge.py", 1000, "--seed", "101")
