#!/bin/bash
# round-2 GPU check U (2 GPUs): the driver's N=2 invocation with the final code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2u_c3_g2.json 2> $O/r2u_c3_g2.err; echo "bench g2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2u_c3_g2.json").read()); r=d["roofline"]
    print("c3_g2 value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "dtype", d["dtype"], "kernel_ms", round(r["kernel_ms_per_step"],4), "frac", round(r["frac"],3), "parity", d.get("parity_checked"), "fb", d.get("exact_fallback_queries"), d.get("per_rank"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks",{}).get("reasons"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2u_c3_g2.err").read()[-3000:])
PY
