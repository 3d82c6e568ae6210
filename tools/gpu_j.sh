#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/dev_decim.py > gpurun_out/j_decim.log 2>&1; cat gpurun_out/j_decim.log | cut -c1-1500
B200ICP_DBG_TAIL=1 timeout 300 python tools/dev_decim.py 2>&1 | grep "dbg tail" | head -5
