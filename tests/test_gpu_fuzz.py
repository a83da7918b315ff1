"""Randomised GPU checks (seeded): the scoring kernels against the fp32 CUDA-core path, and the whole update_kv chain
against its own invariants (indices == stable top-k of the returned scores, gathers bit-exact, TSP set exact)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *args], capture_output=True, text=True,
                       timeout=280, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and "fuzz done" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("seed", ["3", "4"])
def test_scoring_kernels_fuzz(seed):
    _run("fuzz_score.py", "50", seed)


def test_scoring_kernels_fuzz_multi_phase():
    _run("fuzz_score.py", "40", "5", env={"FKV_TC_PPP": "1"})


@pytest.mark.parametrize("seed", ["6", "7"])
def test_update_kv_chain_fuzz(seed):
    _run("fuzz_compress.py", "50", seed)
