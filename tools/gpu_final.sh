#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err
python tools/summarize_bench.py gpurun_out/r02_bench_final3.json 2>/dev/null | head -4
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_final3.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','mean_outer_iterations')})
print('e2e', d['e2e']['value'], 'full', d['e2e_full_module']['registrations_per_s'], 'dec', d['e2e_decimated_1m']['registrations_per_s'], 'C4', d['batch_lc']['registrations_per_s'], d['batch_lc']['outer_iterations_total'], 'C3', d['scan_to_map']['registrations_per_s'], 'cpu', d['cpu_baseline']['value'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm3.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_reference_arm3.json') if l.startswith('{')][-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
