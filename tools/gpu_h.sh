#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
for g in 1 0; do echo "== GRAPH=$g"; B200ICP_GRAPH=$g timeout 300 python tools/dev_e2e.py 2>&1 | cut -c1-900; done > gpurun_out/h_e2e.log 2>&1
cat gpurun_out/h_e2e.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
python tools/summarize_bench.py gpurun_out/h_bench.json 2>/dev/null | head -8 || tail -20 gpurun_out/h_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/h_bench.json') if l.startswith('{')][-1]); print(json.dumps(d['e2e'])[:700])"
grep "module sections" gpurun_out/h_bench.err | cut -c1-600
