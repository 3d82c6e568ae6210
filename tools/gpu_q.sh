#!/bin/bash
# search-stage variants -- parity of the search tests under cta, then A/B timings (same box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/q_ab.log gpurun_out/q_pytest.log
B200ICP_SEARCH=cta timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_match.py tests/test_gpu_align.py tests/test_gpu_multi.py -m gpu -x -q >> gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -12 gpurun_out/q_pytest.log | cut -c1-200
for v in walk cta; do
  echo "== B200ICP_SEARCH=$v" >> gpurun_out/q_ab.log
  B200ICP_SEARCH=$v timeout 300 python tools/dev_align_time.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_SEARCH=$v REPS=5 B200ICP_PROFILE=1 timeout 300 python tools/dev_profile2.py >> gpurun_out/q_ab.log 2>&1
  B200ICP_SEARCH=$v B200ICP_DBG_ITEMS=1 timeout 300 python tools/dev_items.py >> gpurun_out/q_ab.log 2>&1
done
cat gpurun_out/q_ab.log | cut -c1-300
