#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python tools/summarize_bench.py gpurun_out/m_bench.json 2>/dev/null | grep -v "^knn\|^accuracy" || tail -20 gpurun_out/m_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/m_bench.json') if l.startswith('{')][-1]); print({k:d['e2e'][k] for k in ('value','sync_feed_value')}); print(json.dumps(d.get('batch_lc')))"
grep -E "section .* failed" gpurun_out/m_bench.err
