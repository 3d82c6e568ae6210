#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=$PWD/mola-fe-lidar_b200/lib
for lib in libb200icp var_minb5 var_minb6s; do
  echo "== $lib"
  B200ICP_LIB=$L/$lib.so PAIRS=256 timeout 300 python tools/dev_c4.py
  B200ICP_LIB=$L/$lib.so PAIRS=256 timeout 300 python tools/dev_c4.py
done
