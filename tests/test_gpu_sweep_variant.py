"""The alternative search stage (tile sweep, sweep_search.cuh; selected with
B200ICP_SEARCH=sweep, read once per process) must give the same bits as the
default per-lane walk: the kNN / matcher / registration parity tests are re-run
in a child process with the switch set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_parity_suite_with_tile_sweep_search():
    env = dict(os.environ, B200ICP_SEARCH="sweep")
    files = ["tests/test_gpu_knn.py", "tests/test_gpu_match.py", "tests/test_gpu_align.py", "tests/test_gpu_multi.py"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files,
                       cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
