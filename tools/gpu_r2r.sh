#!/bin/bash
# round-2 GPU check R (2 GPUs): sharded finish repairs every outstanding deferred search
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > $O/r2r_multi2.log 2>&1; echo "multi2 rc=$?"
grep "multi-gpu ok" $O/r2r_multi2.log; grep -i "error\|assert\|Traceback" $O/r2r_multi2.log | head -10
