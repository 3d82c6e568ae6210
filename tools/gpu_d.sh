#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for b in 0 96; do
  echo "== DBG items budget $b"
  B200ICP_DBG_ITEMS=1 B200ICP_BUDGET=$b timeout 300 python tools/dev_items.py 2>&1 | grep -v "^\[dbg items\] #"
done > gpurun_out/d_items.log 2>&1
cat gpurun_out/d_items.log
