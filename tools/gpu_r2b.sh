#!/bin/bash
# round-2 GPU check B: private candidate segments, fused tails, latency form, full-size parity, launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_masks.py tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_install_real.py -x -q > $O/r2b_tests.log 2>&1; echo "tests rc=$?"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > $O/r2b_fullsize.log 2>&1; echo "fullsize rc=$?"
for w in c5 c3 c2 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > $O/r2b_$w.json 2> $O/r2b_$w.err; echo "$w rc=$?"
done
timeout 300 python bench.py --workload s1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > $O/r2b_s1.json 2> $O/r2b_s1.err; echo "s1 rc=$?"
timeout 300 python bench.py --workload s8 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > $O/r2b_s8.json 2> $O/r2b_s8.err; echo "s8 rc=$?"
timeout 200 python tools/latency_probe.py > $O/r2b_latency.log 2>&1; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2b_bvb.json > $O/r2b_bvb.log 2>&1; echo "bvb rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2b_launches_c3.csv python bench.py --workload c3 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2b_ncu_c3.err; echo "ncu c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2b_launches_c5.csv python bench.py --workload c5 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2b_ncu_c5.err; echo "ncu c5 rc=$?"
tail -n 4 $O/r2b_tests.log; tail -n 4 $O/r2b_fullsize.log; cat $O/r2b_latency.log; grep -A5 "B200" $O/r2b_bvb.log | grep -E "B200|median"
for f in c5 c3 c2 c1 s1 s8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"))
except Exception as e:
    print("$f failed", e)
PY
done
