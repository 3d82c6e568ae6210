#!/bin/bash
# round-2 dev run A: parity suite + search timing probes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
B200ICP_DBG_ITEMS=1 timeout 300 python tools/dev_items.py > gpurun_out/a_items.log 2>&1
B200ICP_SEED=1 timeout 300 python tools/dev_align_time.py > gpurun_out/a_align_seed1.log 2>&1
B200ICP_SEED=0 timeout 300 python tools/dev_align_time.py > gpurun_out/a_align_seed0.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
cat gpurun_out/a_items.log gpurun_out/a_align_seed1.log gpurun_out/a_align_seed0.log | grep -v "^\[dbg items\] #"
python tools/summarize_bench.py gpurun_out/a_bench.json 2>/dev/null || head -c 1500 gpurun_out/a_bench.json
