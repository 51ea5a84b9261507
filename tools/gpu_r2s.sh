#!/bin/bash
# round-2 GPU check S (1 GPU): final verification after the latency work (slot watching, rank selection, grid bound)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_masks.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "single_launch or predicate_ties or batched_equals or multi_pass or row_mask_on_every_kernel_path" > $O/r2s_racecheck_scan.log 2>&1; echo "racecheck scan rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_masks.py -m gpu -x -q \
  -k "predicate or single_launch or fold" > $O/r2s_memcheck_scan.log 2>&1; echo "memcheck scan rc=$?"
timeout 200 python tools/latency_probe.py > $O/r2s_latency.log 2> $O/r2s_latency.err; echo "latency rc=$?"
TAV_TRACE=1 timeout 200 python tools/latency_probe.py > /dev/null 2> $O/r2s_trace.log; echo "trace rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2s_bvb.json > $O/r2s_bvb.log 2>&1; echo "bvb rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2s_default.json 2> $O/r2s_default.err; echo "default rc=$?"
tail -n 5 $O/r2s_pytest_gpu.log; tail -n 1 $O/r2s_smoke.log
for f in racecheck_scan memcheck_scan; do grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/r2s_$f.log | tail -2; done
cut -c1-330 $O/r2s_latency.log
grep "tav trace" $O/r2s_trace.log | awk '{g=$4; c[g]++; if (c[g]<=1) print}' | cut -c1-260
grep -A5 "B200" $O/r2s_bvb.log | grep -E "B200|median"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2s_default.json").read())
    r=d["roofline"]
    print("default", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), "sus", (r.get("sustained") or {}).get("frac"), "parity", d.get("parity_checked"))
    for k,v in (d.get("secondary") or {}).items():
        print("   sec", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"]), "e2e_ms", round(v["e2e"]["ms_per_step"],4), v["roofline"]["bound"], round(v["roofline"]["frac"],3), "parity", v.get("parity_checked"))
except Exception as e:
    print("default failed", e); print(open("gpurun_out/r2s_default.err").read()[-2000:])
PY
