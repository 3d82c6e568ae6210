#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -15 gpurun_out/e_pytest.log
timeout 200 python tools/dev_align_time.py > gpurun_out/e_align.log 2>&1; cat gpurun_out/e_align.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
python tools/summarize_bench.py gpurun_out/e_bench.json 2>/dev/null || (head -c 600 gpurun_out/e_bench.json; tail -5 gpurun_out/e_bench.err)
