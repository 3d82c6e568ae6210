#!/bin/bash
# round-2 dev run B: parity with the cooperative heavy pass + budget / variant timing matrix
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
L=mola-fe-lidar_b200/lib
: > gpurun_out/b_matrix.log
for lib in libb200icp var_minb4 var_minb4p16 var_plain1000; do
  for b in 0 32 64 96 160 256; do
    echo "== lib $lib budget $b" >> gpurun_out/b_matrix.log
    B200ICP_LIB=$PWD/$L/$lib.so B200ICP_BUDGET=$b timeout 200 python tools/dev_items.py >> gpurun_out/b_matrix.log 2>&1
  done
done
for lib in libb200icp var_minb4 var_minb4p16; do
  for b in 0 64 96 160; do
    echo "== align lib $lib budget $b" >> gpurun_out/b_matrix.log
    B200ICP_LIB=$PWD/$L/$lib.so B200ICP_BUDGET=$b timeout 200 python tools/dev_align_time.py >> gpurun_out/b_matrix.log 2>&1
  done
done
cat gpurun_out/b_matrix.log
