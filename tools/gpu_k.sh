#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -25 gpurun_out/k_pytest.log
timeout 300 python tools/dev_decim.py > gpurun_out/k_decim.log 2>&1; cat gpurun_out/k_decim.log | cut -c1-1200
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
python tools/summarize_bench.py gpurun_out/k_bench.json 2>/dev/null | head -8 || tail -20 gpurun_out/k_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/k_bench.json') if l.startswith('{')][-1]); print(json.dumps(d['e2e'])[:300]); print({k:d['e2e'][k] for k in ('value','sync_feed_value')})"
