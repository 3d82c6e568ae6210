#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=mola-fe-lidar_b200/lib
: > gpurun_out/c_matrix.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/c_matrix.log
for lib in libb200icp var_minb4 var_minb4p16 var_minb4p1000; do
  for b in 0 96 256; do
    echo "== align lib $lib budget $b" >> gpurun_out/c_matrix.log
    B200ICP_LIB=$PWD/$L/$lib.so B200ICP_BUDGET=$b timeout 200 python tools/dev_align_time.py >> gpurun_out/c_matrix.log 2>&1
  done
done
echo "== align lib var_minb4 budget 0 seeds off" >> gpurun_out/c_matrix.log
B200ICP_LIB=$PWD/$L/var_minb4.so B200ICP_BUDGET=0 B200ICP_SEED=0 timeout 200 python tools/dev_align_time.py >> gpurun_out/c_matrix.log 2>&1
cat gpurun_out/c_matrix.log
# ncu: the two search kernels of one registration, default lib, budget 96
B200ICP_BUDGET=96 timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_ -s 2 -c 4 -f -o gpurun_out/r02a_search python tools/dev_profile.py > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
