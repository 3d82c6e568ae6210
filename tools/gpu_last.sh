#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/last_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/last_pytest.log
tail -4 gpurun_out/last_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
