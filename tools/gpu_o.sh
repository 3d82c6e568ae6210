#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_sharded_native.py tests/test_gpu_two_process.py -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -40 gpurun_out/o_pytest.log
