#!/bin/bash
# round-2 GPU check K (1 GPU): the driver's sequence (pytest -m gpu, smoke, reference arm, default bench) + ncu evidence
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2k_reference.json 2> $O/r2k_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2k_default.json 2> $O/r2k_default.err; echo "default rc=$?"
timeout 600 python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline > $O/r2k_c3_40.json 2> $O/r2k_c3_40.err; echo "c3x40 rc=$?"
for w in s1 s8 c2f32 c5f32; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --sustain-seconds 0 > $O/r2k_$w.json 2> $O/r2k_$w.err; echo "$w rc=$?"
done
timeout 200 python tools/latency_probe.py > $O/r2k_latency.log 2>&1; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2k_bvb.json > $O/r2k_bvb.log 2>&1; echo "bvb rc=$?"
B="--no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tav|scan|select|finalize|merge|prep' -c 24 --csv --log-file $O/r2k_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_l3.err; echo "launch list c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tav|scan|select|finalize|merge|prep' -c 24 --csv --log-file $O/r2k_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_l5.err; echo "launch list c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_ts_main -s 3 -c 1 -o $O/r2k_prof_c3_main python bench.py --workload c3 --steps 3 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_c3.err; echo "ncu c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mma_ts_main -s 3 -c 1 -o $O/r2k_prof_c5_main python bench.py --workload c5 --steps 3 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_c5.err; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_rows_kernel -s 2 -c 1 -o $O/r2k_prof_s1_scan python bench.py --workload s1 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_s1.err; echo "ncu s1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_rows_kernel -s 2 -c 1 -o $O/r2k_prof_s8_scan python bench.py --workload s8 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2k_ncu_s8.err; echo "ncu s8 rc=$?"
tail -n 6 $O/r2k_pytest_gpu.log; tail -n 1 $O/r2k_smoke.log; cat $O/r2k_latency.log; grep -A5 "B200\|CPU" $O/r2k_bvb.log | grep -E "B200|CPU|median"
for f in default c3_40 s1 s8 c2f32 c5f32; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2k_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"), (r.get("sustained") or {}).get("sm_mhz"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", d.get("parity_checked"))
    for k,v in (d.get("secondary") or {}).items():
        print("   sec", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"]), v["roofline"]["bound"], round(v["roofline"]["frac"],3), "cpu", (v.get("cpu_baseline") or {}).get("value"), "parity", v.get("parity_checked"))
except Exception as e:
    print("$f failed", e)
PY
done
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2k_reference.json").read()); print("reference", d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
except Exception as e: print("ref failed", e)
PY
