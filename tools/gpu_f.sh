#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -12 gpurun_out/f_pytest.log
B200ICP_DBG_TAIL=1 timeout 200 python tools/dev_align_time.py 2>&1 | sort | uniq -c | sort -rn | head -12
