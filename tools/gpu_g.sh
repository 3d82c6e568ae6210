#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/g_ref.json 2> gpurun_out/g_ref.err
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
tail -c 1500 gpurun_out/g_ref.json; echo
python tools/summarize_bench.py gpurun_out/g_bench.json 2>/dev/null || (head -c 600 gpurun_out/g_bench.json; tail -20 gpurun_out/g_bench.err)
grep -E "C3 map|section .* failed" gpurun_out/g_bench.err
