#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for g in 1 0; do echo "== GRAPH=$g"; B200ICP_GRAPH=$g timeout 300 python tools/dev_async.py 2>&1 | cut -c1-700; done > gpurun_out/i_async.log 2>&1
cat gpurun_out/i_async.log
