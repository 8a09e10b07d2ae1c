"""Bounded runs of the differential fuzzers (tests/fuzz_host.py, tests/fuzz_gpu.py): a fixed seed and a fixed number of
cases each (the same cases every time), a few seconds.
The long runs and what they found: DESIGN.md section 5.1; the findings themselves are pinned in tests/golden/frames_fuzz.npz."""
import subprocess
import sys

import pytest

from helpers import ROOT


def run(tool, *args):
    p = subprocess.run([sys.executable, str(ROOT / "tests" / tool), *args], capture_output=True, text=True, timeout=600)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:]
    return p.returncode, tail, p.stdout


def test_host_fuzz_oracle_vs_parser():
    """reference (when it is built here) vs oracle vs the product's parser + planner on damaged streams"""
    rc, tail, out = run("fuzz_host.py", "--cases", "400", "--seed", "11", "--families", "ht_sets,layered,plain,packed,tile_styles")
    assert rc == 0 and " 0 differ" in tail, out[-2000:]


@pytest.mark.gpu
def test_gpu_fuzz_product_vs_oracle():
    """the product through the C ABI (all kernels) vs the oracle on damaged streams, HT and Part-1"""
    rc, tail, out = run("fuzz_gpu.py", "--cases", "6000", "--seed", "12")
    assert rc == 0 and " 0 differ" in tail, out[-2000:]
    rc, tail, out = run("fuzz_gpu.py", "--cases", "2000", "--seed", "13", "--families", "part1,subsampled")
    assert rc == 0 and " 0 differ" in tail, out[-2000:]
