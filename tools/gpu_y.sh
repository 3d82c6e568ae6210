#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest.log
tail -6 gpurun_out/y_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/y_bench.json') if l.startswith('{')][-1])
print('auto value %.0f e2e %.0f full %.0f C4 %.0f C3 %.0f'%(d['value'], d['e2e']['value'], d['e2e_full_module']['registrations_per_s'], d['batch_lc']['registrations_per_s'], d['scan_to_map']['registrations_per_s']))
PY
