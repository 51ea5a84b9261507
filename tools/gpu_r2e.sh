#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_masks.py tests/test_gpu_parity.py tests/test_install_real.py -x -q > $O/r2e_tests.log 2>&1; echo "tests rc=$?"
for w in c5 c3 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-parity > $O/r2e_$w.json 2> $O/r2e_$w.err; echo "$w rc=$?"
done
timeout 200 python tools/latency_probe.py > $O/r2e_latency.log 2>&1; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2e_bvb.json > $O/r2e_bvb.log 2>&1; echo "bvb rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tav|scan|select|finalize|merge|prep' -c 40 --csv --log-file $O/r2e_launches_c5.csv python bench.py --workload c5 --steps 3 --warmup 1 --no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0 > /dev/null 2> $O/r2e_ncu_c5.err; echo "ncu c5 rc=$?"
tail -n 6 $O/r2e_tests.log; cat $O/r2e_latency.log; grep -A5 "B200" $O/r2e_bvb.log | grep -E "B200|median"
for f in c5 c3 c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"), (r.get("sustained") or {}).get("sm_mhz"), (r.get("sustained") or {}).get("kernel_ms"))
except Exception as e:
    print("$f failed", e)
PY
done
python - <<'PY'
import csv,collections
rows=[l for l in open('gpurun_out/r2e_launches_c5.csv') if l.startswith('"')]
d=collections.OrderedDict()
for x in csv.DictReader(rows):
    key=x['Kernel Name'].split('(')[0][-40:]+" grid"+x['Grid Size']
    d.setdefault(key,[]).append(float(x['Metric Value'])/1e3)
for k,v in d.items(): print("  ",k, len(v), 'avg us', round(sum(v)/len(v),1))
PY
