#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -25 gpurun_out/l_pytest.log
timeout 300 python tools/dev_decim.py > gpurun_out/l_decim.log 2>&1; cat gpurun_out/l_decim.log | cut -c1-900
