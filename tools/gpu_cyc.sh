#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_lidar_odometry.py tests/test_gpu_solvers.py tests/test_gpu_sharded_native.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/dev_decim.py 2>&1 | cut -c1-700
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/cyc_bench.json 2> gpurun_out/cyc_bench.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/cyc_bench.json') if l.startswith('{')][-1])
print('value %.0f e2e %.0f full %.0f dec %.0f'%(d['value'], d['e2e']['value'], d['e2e_full_module']['registrations_per_s'], d['e2e_decimated_1m']['registrations_per_s']))
PY
grep "module sections" gpurun_out/cyc_bench.err | sed -n 4p | cut -c1-500
