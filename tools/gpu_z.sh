#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
python tools/summarize_bench.py gpurun_out/r02_bench_final2.json 2>/dev/null | head -6 || tail -20 gpurun_out/r02_bench_final2.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_final2.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','cuda_graph_replays_timed_region','profiled_pass_ms_per_step','mean_outer_iterations')})
print('e2e', d['e2e']['value'], 'full', d['e2e_full_module']['registrations_per_s'], 'dec', d['e2e_decimated_1m']['registrations_per_s'], 'C4', d['batch_lc']['registrations_per_s'], 'cpu', d['cpu_baseline']['value'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 600
