"""A short slice of the randomized sweeps under scripts/ (fuzz_parity.py, fuzz_pool.py, fuzz_merge.py) inside the GPU
suite: seeded, time-boxed, run as the scripts themselves so that a replay (`--cases`) of anything they report uses
the same code.  The long runs are recorded in profiles/r02c_fuzz.txt."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-15:])
    assert res.returncode == 0, f"{script} {' '.join(args)} reported mismatches:\n{tail}"
    last = res.stdout.strip().splitlines()[-1]
    assert " 0 failing" in last, last
    cases = int(last.split(":")[1].split("cases")[0])
    assert cases >= 20, f"only {cases} cases ran: {last}"


def test_search_whatever_the_planner_picks_equals_the_float64_scan():
    _run("fuzz_parity.py", "--seconds", "12", "--seed", "101")


def test_search_big_shapes_equal_the_float64_scan():
    _run("fuzz_parity.py", "--seconds", "10", "--seed", "102", "--big")


def test_pool_equals_the_float64_formula():
    _run("fuzz_pool.py", "--seconds", "4", "--seed", "101")


def test_merge_equals_a_stable_sort():
    _run("fuzz_merge.py", "--seconds", "3", "--seed", "101")
