#!/bin/bash
# round 2b: item sweep -- parity of the search tests, then A/B timings (same box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/q_ab.log
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_match.py tests/test_gpu_align.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -15 gpurun_out/q_pytest.log
for v in 1 0; do
  echo "== B200ICP_ITEM_SWEEP=$v" >> gpurun_out/q_ab.log
  B200ICP_ITEM_SWEEP=$v timeout 300 python tools/dev_align_time.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_ITEM_SWEEP=$v timeout 300 python tools/dev_items.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_ITEM_SWEEP=$v B200ICP_DBG_ITEMS=1 timeout 300 python tools/dev_items.py >> gpurun_out/q_ab.log 2>&1
done
echo "== phases variant" >> gpurun_out/q_ab.log
B200ICP_LIB=$PWD/mola-fe-lidar_b200/lib/var_phases.so timeout 300 python tools/dev_items.py >> gpurun_out/q_ab.log 2>&1
cat gpurun_out/q_ab.log | cut -c1-400
