#!/bin/bash
# round 2b: search-stage variants -- parity of the search tests under each, then A/B timings (same box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/q_ab.log gpurun_out/q_pytest.log
for v in "item 0" "item 1"; do
  set -- $v
  echo "== pytest B200ICP_SEARCH=$1 B200ICP_TMA=$2" >> gpurun_out/q_pytest.log
  B200ICP_SEARCH=$1 B200ICP_TMA=$2 timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_match.py tests/test_gpu_align.py -m gpu -x -q >> gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
done
grep -E "passed|failed|rc=|==" gpurun_out/q_pytest.log
for v in "walk 0" "item 0" "item 1"; do
  set -- $v
  echo "== B200ICP_SEARCH=$1 B200ICP_TMA=$2" >> gpurun_out/q_ab.log
  B200ICP_SEARCH=$1 B200ICP_TMA=$2 timeout 300 python tools/dev_align_time.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_SEARCH=$1 B200ICP_TMA=$2 REPS=5 B200ICP_PROFILE=1 timeout 300 python tools/dev_profile2.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_SEARCH=$1 B200ICP_TMA=$2 B200ICP_DBG_ITEMS=1 timeout 300 python tools/dev_items.py >> gpurun_out/q_ab.log 2>&1
done
cat gpurun_out/q_ab.log | cut -c1-300
B200ICP_SEARCH=item B200ICP_TMA=1 B200ICP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_item -s 1 -c 1 -f -o gpurun_out/r02d_search_tma python tools/dev_profile2.py > gpurun_out/q_ncu.log 2>&1
tail -2 gpurun_out/q_ncu.log
