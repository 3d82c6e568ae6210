#!/bin/bash
# round 2 final evidence, 1 GPU: bench (both arms), launch list, ncu of the shipped search + fit kernels, sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/t_smi.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python tools/summarize_bench.py gpurun_out/r02_bench_final.json 2>/dev/null | head -12 || tail -20 gpurun_out/r02_bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
tail -c 1500 gpurun_out/r02_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/t_ncu_launch.log 2>&1
B200ICP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tile_kernel -s 1 -c 2 -f -o gpurun_out/r02_search_walk python tools/dev_profile2.py > gpurun_out/t_ncu_search.log 2>&1
B200ICP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_plane_kernel -s 1 -c 1 -f -o gpurun_out/r02_fit python tools/dev_profile2.py > gpurun_out/t_ncu_fit.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" >> gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_knn.py tests/test_gpu_edges_planes.py -m gpu -x -q -k "not lidar_scan and not module" > gpurun_out/r02_sanitizer_memcheck_knn.log 2>&1; echo "memcheck knn rc=$?" >> gpurun_out/r02_sanitizer_memcheck_knn.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" >> gpurun_out/r02_sanitizer_racecheck_smoke.log
tail -4 gpurun_out/r02_sanitizer_*.log
