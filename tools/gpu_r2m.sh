#!/bin/bash
# round-2 GPU check M (8 GPUs): final sharded numbers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > $O/r2m_multi8.log 2>&1; echo "multi8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2m_c3_g8.json 2> $O/r2m_c3_g8.err; echo "bench g8 rc=$?"
grep "multi-gpu ok" $O/r2m_multi8.log | tail -n 9; tail -n 3 $O/r2m_multi8.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2m_c3_g8.json").read())
    r=d["roofline"]
    print("c3_g8", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "kernel_ms", round(r["kernel_ms_per_step"],4), "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "parity", d["parity_checked"], "fb", d["exact_fallback_queries"], "sus", (r.get("sustained") or {}).get("frac"), (r.get("sustained") or {}).get("value"))
    for k,v in (d.get("secondary") or {}).items():
        rr=v["roofline"]
        print("   sec", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"]), rr["bound"], round(rr["frac"],3), "kernel_ms", round(rr["kernel_ms_per_step"],4), rr["per_step_ms_by_kernel_kind"], "parity", v.get("parity_checked"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2m_c3_g8.err").read()[-2500:])
PY
