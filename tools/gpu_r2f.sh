#!/bin/bash
# round-2 GPU check F (2 GPUs): sharded bit-identity with the peer exchange, sharded bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2f_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > $O/r2f_multi2.log 2>&1; echo "multi2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_install_real.py -x -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2f_c3_g2.json 2> $O/r2f_c3_g2.err; echo "bench g2 rc=$?"
tail -n 12 $O/r2f_multi2.log; tail -n 4 $O/r2f_tests.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2f_c3_g2.json").read())
    r=d["roofline"]
    print("c3 g2", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(r["kernel_ms_per_step"],4), "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "parity", d["parity_checked"], "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"))
except Exception as e:
    print("bench g2 failed", e); print(open("gpurun_out/r2f_c3_g2.err").read()[-3000:])
PY
