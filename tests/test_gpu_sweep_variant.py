"""The alternative search stages -- the radius-wide tile sweep (sweep_search.cuh, B200ICP_SEARCH=sweep), the item
sweep (item_sweep.cuh, B200ICP_SEARCH=item) and its cp.async.bulk + mbarrier staging (B200ICP_TMA=1); the switches are
read once per process -- must give the same bits as the default per-lane walk: the kNN / matcher / registration
parity tests are re-run in a child process with the switch set."""
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
@pytest.mark.parametrize("tma", ["0", "1"])
def test_parity_suite_with_item_sweep_search(tma):
    _run({"B200ICP_SEARCH": "item", "B200ICP_TMA": tma},
         ["tests/test_gpu_knn.py", "tests/test_gpu_match.py", "tests/test_gpu_align.py", "tests/test_gpu_multi.py"])
