#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lidar_odometry.py tests/test_gpu_align.py tests/test_gpu_voxel.py tests/test_gpu_edges_planes.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/u_bench$i.json 2> gpurun_out/u_bench$i.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/u_bench$i.json') if l.startswith('{')][-1])
print('run $i value %.0f e2e %.0f full %.0f dec %.0f'%(d['value'], d['e2e']['value'], d['e2e_full_module']['registrations_per_s'], d['e2e_decimated_1m']['registrations_per_s']))
PY
grep "module sections" gpurun_out/u_bench$i.err | sed -n 3p | cut -c1-420
done
