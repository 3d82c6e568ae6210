#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/p_bench_n2.json 2> gpurun_out/p_bench_n2.err
python tools/summarize_bench.py gpurun_out/p_bench_n2.json 2>/dev/null | head -5 || tail -30 gpurun_out/p_bench_n2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/p_bench_n2.json') if l.startswith('{')][-1]); print(json.dumps(d.get('sharded_native'),indent=0)[:2500]); print(d.get('extras_error')); print(json.dumps(d.get('batch_lc'))[:300])"
