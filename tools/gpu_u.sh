#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/u_bench1.json 2> gpurun_out/u_bench1.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/u_bench1.json') if l.startswith('{')][-1])
print('value %.0f (%.3f ms) e2e %.0f windows %s full %.0f dec %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['windows_ms_rank0'], d['e2e_full_module']['registrations_per_s'], d['e2e_decimated_1m']['registrations_per_s']))
PY
tail -3 gpurun_out/u_bench1.err | cut -c1-200
