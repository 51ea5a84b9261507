#!/bin/bash
# round-2 GPU check T (1 GPU): sweep of the small-k admission target on c5 (50k x 384 bf16, B=1000, k=5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
B="--workload c5 --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --sustain-seconds 0"
for T in 32 48 64 96 128 256; do
  TAV_SMALLK_TARGET=$T timeout 200 python bench.py $B > $O/r2t_c5_$T.json 2> $O/r2t_c5_$T.err; echo "c5 target $T rc=$?"
done
TAV_SMALLK_TARGET=64 timeout 300 python -m pytest tests/test_gpu_mma.py -x -q -m gpu -k "small_k or in_register or many_chunks" > $O/r2t_tests64.log 2>&1; echo "tests64 rc=$?"
TAV_SMALLK_TARGET=128 timeout 300 python -m pytest tests/test_gpu_mma.py -x -q -m gpu -k "small_k or in_register or many_chunks" > $O/r2t_tests128.log 2>&1; echo "tests128 rc=$?"
tail -n 1 $O/r2t_tests64.log $O/r2t_tests128.log
python - <<'PY'
import json
for t in (32,48,64,96,128,256):
    try:
        d=json.loads(open(f"gpurun_out/r2t_c5_{t}.json").read()); r=d["roofline"]
        print("target", t, "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "search", round(r.get("search_ms_per_step_same_pass",0),4), {k:round(v,4) for k,v in r["per_step_ms_by_kernel_kind"].items()}, "parity", d.get("parity_checked"), "fb", d.get("exact_fallback_queries"))
    except Exception as e:
        print("target", t, "failed", e)
PY
