#!/bin/bash
# bench at N GPUs (N = $1), both arms
N=$1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python tools/summarize_bench.py gpurun_out/r02_bench_n$N.json 2>/dev/null | head -5 || tail -30 gpurun_out/r02_bench_n$N.err
timeout 600 python -m pytest tests/test_gpu_two_process.py tests/test_gpu_sharded_native.py -m gpu -x -q 2>&1 | tail -3
