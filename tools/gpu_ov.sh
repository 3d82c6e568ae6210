#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for g in 1 0; do for ov in 0 1; do
B200ICP_GRAPH=$g B200ICP_OVERLAP=$ov timeout 120 python tools/dev_overlap.py
done; done
grep "module sections" gpurun_out/ov_bench0.err 2>/dev/null | sed -n 3p | cut -c1-500
