#!/bin/bash
# round-2 GPU check D: tightened TS issue loop, histogram select in the fused scan, ncu full of the c5 / c3 MAIN kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_masks.py tests/test_gpu_parity.py -x -q > $O/r2d_tests.log 2>&1; echo "tests rc=$?"
for w in c5 c3 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > $O/r2d_$w.json 2> $O/r2d_$w.err; echo "$w rc=$?"
done
TAV_NO_TS=1 timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > $O/r2d_c3_smem.json 2> $O/r2d_c3_smem.err; echo "c3 smem rc=$?"
timeout 200 python tools/latency_probe.py > $O/r2d_latency.log 2>&1; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2d_bvb.json > $O/r2d_bvb.log 2>&1; echo "bvb rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_ts_main -s 3 -c 1 -o $O/r2d_prof_c5_ts python bench.py --workload c5 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2d_ncu_c5.err; echo "ncu c5 rc=$?"
TAV_NO_TS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mma_topk_kernel<1' -s 3 -c 1 -o $O/r2d_prof_c5_ss python bench.py --workload c5 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2d_ncu_c5ss.err; echo "ncu c5 ss rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_ts_main -s 3 -c 1 -o $O/r2d_prof_c3_ts python bench.py --workload c3 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2d_ncu_c3.err; echo "ncu c3 rc=$?"
tail -n 6 $O/r2d_tests.log; cat $O/r2d_latency.log; grep -A5 "B200" $O/r2d_bvb.log | grep -E "B200|median"
for f in c5 c3 c3_smem c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"), (r.get("sustained") or {}).get("sm_mhz"), (r.get("sustained") or {}).get("kernel_ms"))
except Exception as e:
    print("$f failed", e)
PY
done
