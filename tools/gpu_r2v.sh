#!/bin/bash
# round-2 GPU check V (1 GPU): the driver's round-end sequence (gpu tests, smoke) on the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2v_bvb.json > $O/r2v_bvb.log 2>&1; echo "bvb rc=$?"
tail -n 5 $O/r2v_pytest_gpu.log; tail -n 1 $O/r2v_smoke.log; grep -A5 "B200" $O/r2v_bvb.log | grep -E "B200|median"
