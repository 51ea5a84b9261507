#!/bin/bash
# round-2 GPU check I (2 GPUs): merge with integrated wait/ack, adaptive sampler blocks, scan trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_masks.py tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q > $O/r2i_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2i_c3_g2.json 2> $O/r2i_c3_g2.err; echo "bench g2 rc=$?"
TAV_TRACE=1 timeout 200 python tools/latency_probe.py > $O/r2i_latency.log 2> $O/r2i_trace.log; echo "latency rc=$?"
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-secondary --no-parity > $O/r2i_c1.json 2> $O/r2i_c1.err; echo "c1 rc=$?"
timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > $O/r2i_c5.json 2> $O/r2i_c5.err; echo "c5 rc=$?"
tail -n 5 $O/r2i_tests.log; cat $O/r2i_latency.log; grep "tav trace" $O/r2i_trace.log | head -12
for f in c3_g2 c1 c5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "parity", d["parity_checked"], "fb", d["exact_fallback_queries"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2i_$f.err").read()[-2500:])
PY
done
