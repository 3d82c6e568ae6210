#!/bin/bash
# throughput-bound workloads (C4 batch, big-map kNN, C5) under the walk and under the item sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in walk item; do
B200ICP_SEARCH=$v timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_$v.json 2> gpurun_out/x_bench_$v.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x_bench_$v.json') if l.startswith('{')][-1])
print('$v value %.0f C4 %.0f C3 %.0f'%(d['value'], d['batch_lc']['registrations_per_s'], d['scan_to_map']['registrations_per_s']))
for k in d['knn']: print('  ', k['case'], k['k'], '%.3f ms'%k.get('kernel_ms',-1))
sk=d.get('sharded_knn') or {}
print('  C5', json.dumps(sk)[:400])
PY
done
