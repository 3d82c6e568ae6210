#!/bin/bash
# ncu of the item-sweep search kernel: first (unseeded) and second (seeded) matcher search of one registration
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tile -s 0 -c 2 -f -o gpurun_out/r02b_search python tools/dev_profile.py > gpurun_out/r_ncu.log 2>&1
tail -3 gpurun_out/r_ncu.log
