"""The search stages -- the per-lane walk (tile_search.cuh), the item sweep (item_sweep.cuh) with and without its
cp.async.bulk + mbarrier staging (B200ICP_TMA=1), the radius-wide tile sweep (sweep_search.cuh) -- must give the same
bits.  By default the library picks the walk for single jobs and the item sweep for batched launches; the kNN /
matcher / registration / batch parity tests are re-run in a child process with each stage FORCED for every launch
(B200ICP_SEARCH, read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, files):
    env = dict(os.environ, **extra_env)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files,
                       cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.timeout(900)
def test_parity_suite_with_tile_sweep_search():
    _run({"B200ICP_SEARCH": "sweep"},
         ["tests/test_gpu_knn.py", "tests/test_gpu_match.py", "tests/test_gpu_align.py", "tests/test_gpu_multi.py"])


@pytest.mark.timeout(900)
def test_parity_suite_with_walk_forced():
    _run({"B200ICP_SEARCH": "walk"},
         ["tests/test_gpu_knn.py", "tests/test_gpu_match.py", "tests/test_gpu_align.py", "tests/test_gpu_multi.py"])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("tma", ["0", "1"])
def test_parity_suite_with_item_sweep_search(tma):
    _run({"B200ICP_SEARCH": "item", "B200ICP_TMA": tma},
         ["tests/test_gpu_knn.py", "tests/test_gpu_match.py", "tests/test_gpu_align.py", "tests/test_gpu_multi.py"])
