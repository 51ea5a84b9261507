#!/bin/bash
# round-2 GPU check O: compute-sanitizer over the new kernels (memcheck, racecheck), then the full gpu suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_masks.py -m gpu -x -q \
  -k "predicate or single_launch or fold" > $O/r2o_memcheck_scan.log 2>&1; echo "memcheck scan rc=$?"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mma.py -m gpu -x -q \
  -k "tensor_memory_form and (777 or 20011 or 33000) or small_k_ties or thresholds_on or fallback_when" > $O/r2o_memcheck_mma.log 2>&1; echo "memcheck mma rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_masks.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "single_launch or predicate_ties or batched_equals or multi_pass or row_mask_on_every_kernel_path" > $O/r2o_racecheck_scan.log 2>&1; echo "racecheck scan rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mma.py -m gpu -x -q \
  -k "tensor_memory_form and 777" > $O/r2o_racecheck_mma.log 2>&1; echo "racecheck mma rc=$?"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/r2o_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2o_smoke.log 2>&1; echo "smoke rc=$?"
for f in memcheck_scan memcheck_mma racecheck_scan racecheck_mma; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/r2o_$f.log | tail -3; grep -E "Race reported" $O/r2o_$f.log | sed 's/\[.*//' | sort | uniq -c | head -6; done
tail -n 4 $O/r2o_pytest_gpu.log; tail -n 1 $O/r2o_smoke.log
