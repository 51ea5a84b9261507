#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_masks.py tests/test_gpu_parity.py tests/test_install_real.py tests/test_gpu_hypothesis.py -x -q > $O/r2g_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python __graft_entry__.py > $O/r2g_build.log 2>&1; python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke.log 2>&1; echo "smoke rc=$?"
for w in c5 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > $O/r2g_$w.json 2> $O/r2g_$w.err; echo "$w rc=$?"
done
timeout 200 python tools/latency_probe.py > $O/r2g_latency.log 2>&1; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2g_bvb.json > $O/r2g_bvb.log 2>&1; echo "bvb rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2g_default.json 2> $O/r2g_default.err; echo "default rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2g_reference.json 2> $O/r2g_reference.err; echo "reference rc=$?"
tail -n 5 $O/r2g_tests.log; tail -n 2 $O/r2g_smoke.log; cat $O/r2g_latency.log; grep -A5 "B200" $O/r2g_bvb.log | grep -E "B200|median"
for f in c5 c1 default; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", d.get("parity_checked"))
    for k,v in (d.get("secondary") or {}).items():
        print("   sec", k, round(v["value"]), "e2e", round(v["e2e"]["value"]), v["roofline"]["bound"], round(v["roofline"]["frac"],3), "cpu", (v.get("cpu_baseline") or {}).get("value"), "parity", v.get("parity_checked"))
except Exception as e:
    print("$f failed", e)
PY
done
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2g_reference.json").read()); print("reference", d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["sample"])
except Exception as e: print("ref failed", e)
PY
