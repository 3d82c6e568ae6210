#!/bin/bash
# ncu of the item-sweep search kernel: the matcher searches of one registration from a velocity-model guess
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B200ICP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_item -s 0 -c 3 -f -o gpurun_out/r02c_search python tools/dev_profile2.py > gpurun_out/r_ncu.log 2>&1
tail -3 gpurun_out/r_ncu.log
B200ICP_LIB=$PWD/mola-fe-lidar_b200/lib/var_phases.so KNN=1 timeout 300 python tools/dev_profile2.py 2>&1 | tail -3
