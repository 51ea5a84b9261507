#!/bin/bash
# round-2 GPU check L (2 GPUs): exchange on the group's own stream, scan hand-over tweaks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_masks.py tests/test_gpu_parity.py tests/test_gpu_hypothesis.py tests/test_gpu_multi.py tests/test_vectorbase_api.py -x -q -m gpu > $O/r2l_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2l_c3_g2.json 2> $O/r2l_c3_g2.err; echo "bench g2 rc=$?"
TAV_TRACE=1 timeout 200 python tools/latency_probe.py > $O/r2l_latency.log 2> $O/r2l_trace.log; echo "latency rc=$?"
timeout 200 python tools/benchmark_vectorbase_gpu.py --json $O/r2l_bvb.json > $O/r2l_bvb.log 2>&1; echo "bvb rc=$?"
tail -n 5 $O/r2l_tests.log; cat $O/r2l_latency.log; grep "tav trace" $O/r2l_trace.log | awk 'NR%4==1' | head -6 | cut -c1-260; grep -A5 "B200" $O/r2l_bvb.log | grep -E "B200|median"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2l_c3_g2.json").read())
    r=d["roofline"]
    print("c3_g2", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "parity", d["parity_checked"], "fb", d["exact_fallback_queries"], d.get("per_rank"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2l_c3_g2.err").read()[-2500:])
PY
