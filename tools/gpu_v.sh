#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B200ICP_DBG_TAIL=1 REPS=3 timeout 300 python tools/dev_profile2.py > gpurun_out/v_tail.log 2>&1; tail -5 gpurun_out/v_tail.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" >> gpurun_out/r02_sanitizer_racecheck_smoke.log
tail -6 gpurun_out/r02_sanitizer_racecheck_smoke.log
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_solvers.py -m gpu -x -q 2>&1 | tail -3
