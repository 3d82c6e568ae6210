#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_align.py tests/test_gpu_lidar_odometry.py tests/test_gpu_solvers.py tests/test_gpu_edges_planes.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/u_bench$i.json 2> gpurun_out/u_bench$i.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/u_bench$i.json') if l.startswith('{')][-1])
print('run $i value %.0f (%.3f ms) e2e %.0f windows %s full %.0f dec %.0f q %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['windows_ms_rank0'], d['e2e_full_module']['registrations_per_s'], d['e2e_decimated_1m']['registrations_per_s'], d['accuracy']['mean_quality']))
PY
done
