#!/bin/bash
# round-2 GPU check A: new kernels' parity + first bench numbers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests/test_gpu_mma.py -x -q > gpurun_out/r2a_mma.log 2>&1; echo "mma rc=$?" 
timeout 300 python -m pytest tests/test_gpu_masks.py -x -q > gpurun_out/r2a_masks.log 2>&1; echo "masks rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -x -q > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"
timeout 300 python bench.py --workload c5 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2a_c5.json 2> gpurun_out/r2a_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2a_c3.json 2> gpurun_out/r2a_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2a_c1.json 2> gpurun_out/r2a_c1.err; echo "c1 rc=$?"
tail -3 gpurun_out/r2a_mma.log gpurun_out/r2a_masks.log gpurun_out/r2a_parity.log
for f in c5 c3 c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", r["kernel_ms_per_step"], "frac", round(r["frac"],3), r["per_step_ms_by_kernel_kind"], "parity", d["parity_checked"], "fb", d["exact_fallback_queries"], (r.get("sustained") or {}).get("frac"))
except Exception as e:
    print("$f failed", e)
PY
done
