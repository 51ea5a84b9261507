#!/bin/bash
# round-2 GPU check Q (1 GPU): rank-selection compaction / merge in the row-scan kernel; latency sweep of the grid bound
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_masks.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "single_launch or predicate_ties or batched_equals or multi_pass or row_mask_on_every_kernel_path" > $O/r2q_racecheck_scan.log 2>&1; echo "racecheck scan rc=$?"
for S in 8192 4096 2048 1024 512; do
  TAV_SCAN1_SURVIVORS=$S timeout 200 python tools/latency_probe.py > $O/r2q_latency_$S.log 2> $O/r2q_latency_$S.err; echo "latency $S rc=$?"
done
TAV_TRACE=1 timeout 200 python tools/latency_probe.py > /dev/null 2> $O/r2q_trace.log; echo "trace rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2q_bvb.json > $O/r2q_bvb.log 2>&1; echo "bvb rc=$?"
tail -n 5 $O/r2q_pytest_gpu.log
grep -E "RACECHECK SUMMARY|passed|failed" $O/r2q_racecheck_scan.log | tail -2
for S in 8192 4096 2048 1024 512; do echo "== survivors $S"; cat $O/r2q_latency_$S.log | cut -c1-330; done
grep "tav trace" $O/r2q_trace.log | awk '{g=$4; if (g!=last) {print; last=g}}' | head -8 | cut -c1-260
grep -A5 "B200" $O/r2q_bvb.log | grep -E "B200|median"
