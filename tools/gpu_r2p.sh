#!/bin/bash
# round-2 GPU check P (2 GPUs): light timing mode in the timed region
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu > $O/r2p_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2p_c3_g2.json 2> $O/r2p_c3_g2.err; echo "bench g2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2p_default.json 2> $O/r2p_default.err; echo "default rc=$?"
tail -n 3 $O/r2p_tests.log
for f in c3_g2 default; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_$f.json").read())
    r=d["roofline"]
    print("$f", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(r["kernel_ms_per_step"],4), "search_same_pass", r.get("search_ms_per_step_same_pass"), r["bound"], "frac", round(r["frac"],3), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "sus", (r.get("sustained") or {}).get("frac"), "parity", d.get("parity_checked"), d.get("per_rank"))
    for k,v in (d.get("secondary") or {}).items():
        print("   sec", k, round(v["value"]), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"]), v["roofline"]["bound"], round(v["roofline"]["frac"],3), v["roofline"]["per_step_ms_by_kernel_kind"], "parity", v.get("parity_checked"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2p_$f.err").read()[-2000:])
PY
done
