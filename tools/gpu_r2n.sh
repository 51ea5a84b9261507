#!/bin/bash
# round-2 GPU check N (1 GPU): final verification of the driver's sequence + launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2n_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2n_default.json 2> $O/r2n_default.err; echo "default rc=$?"
TAV_TRACE=1 timeout 200 python tools/latency_probe.py > $O/r2n_latency.log 2> $O/r2n_trace.log; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2n_bvb.json > $O/r2n_bvb.log 2>&1; echo "bvb rc=$?"
B="--no-secondary --no-cpu-baseline --no-parity --sustain-seconds 0"
K="regex:mma|scan_rows|select_kernel|finalize|merge|prep|publish|split_rows"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 16 --csv --log-file $O/r2n_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2n_ncu_l3.err; echo "launch list c3 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 16 --csv --log-file $O/r2n_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2n_ncu_l5.err; echo "launch list c5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 16 --csv --log-file $O/r2n_launches_c1.csv python bench.py --workload c1 --steps 2 --warmup 1 $B > /dev/null 2> $O/r2n_ncu_l1.err; echo "launch list c1 rc=$?"
tail -n 5 $O/r2n_pytest_gpu.log; tail -n 1 $O/r2n_smoke.log; cat $O/r2n_latency.log; grep "tav trace" $O/r2n_trace.log | awk 'NR%4==1' | head -4 | cut -c1-260; grep -A5 "B200" $O/r2n_bvb.log | grep -E "B200|median"
python - <<'PY'
import json, csv, collections
try:
    d=json.loads(open("gpurun_out/r2n_default.json").read())
    r=d["roofline"]
    print("default", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(r["kernel_ms_per_step"],4), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "sus", (r.get("sustained") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", d.get("parity_checked"))
    for k,v in (d.get("secondary") or {}).items():
        print("   sec", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"]), v["roofline"]["bound"], round(v["roofline"]["frac"],3), "cpu", (v.get("cpu_baseline") or {}).get("value"), "parity", v.get("parity_checked"))
except Exception as e:
    print("default failed", e)
for f in ("c3","c5","c1"):
    try:
        rows=[l for l in open(f'gpurun_out/r2n_launches_{f}.csv') if l.startswith('"')]
        d=collections.OrderedDict()
        for x in csv.DictReader(rows):
            key=x['Kernel Name'].split('(')[0][-44:]+" grid"+x['Grid Size']
            d.setdefault(key,[]).append(float(x['Metric Value'])/1e3)
        print(f)
        for k,v in d.items(): print("  ",k, len(v), 'avg us', round(sum(v)/len(v),1))
    except Exception as e: print(f, "failed", e)
PY
