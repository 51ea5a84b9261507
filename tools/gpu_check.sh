#!/bin/bash
# One gpurun call: GPU tests, smoke, small benches; everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [stage ...]'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
stages="${*:-info tests smoke bench_c1 bench_c3}"
for st in $stages; do
  echo "=== stage $st ($(date +%T))"
  case $st in
    info)
      nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
      nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt ;;
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log ;;
    bench_c1)
      timeout 300 python bench.py --workload c1 --steps 50 --warmup 10 2>gpurun_out/bench_c1.err | tee gpurun_out/bench_c1.json ;;
    bench_c2)
      timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 2>gpurun_out/bench_c2.err | tee gpurun_out/bench_c2.json ;;
    bench_c3)
      timeout 900 python bench.py --steps 40 --warmup 5 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json ;;
    tests_mma)
      timeout 600 python -m pytest tests/test_gpu_mma.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_mma.log ;;
    bench_c5f32)
      timeout 600 python bench.py --workload c5f32 --steps 10 --warmup 3 2>gpurun_out/bench_c5f32.err | tee gpurun_out/bench_c5f32.json ;;
    bench_c2f32)
      timeout 600 python bench.py --workload c2f32 --steps 10 --warmup 3 2>gpurun_out/bench_c2f32.err | tee gpurun_out/bench_c2f32.json ;;
    bench_c5)
      timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 2>gpurun_out/bench_c5.err | tee gpurun_out/bench_c5.json ;;
    ncu_c1)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
        --log-file gpurun_out/launches_c1.csv python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1 ;;
    ncu_c3_list)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tav|mma_topk|threshold|finalize|query_prep|scan_rows|select|merge" -c 60 --csv \
        --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3_list.log 2>&1 ;;
    ncu_c3_full)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mma_topk -s 3 -c 1 \
        -o gpurun_out/prof_c3 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3_full.log 2>&1 ;;
    multi2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/multi_gpu_check.py 2>&1 | tail -15 | tee gpurun_out/multi2.log ;;
    bench_g2)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
        bench.py --gpus 2 --steps 40 --warmup 5 2>gpurun_out/bench_g2.err | tee gpurun_out/bench_g2.json ;;
    multi8)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        tools/multi_gpu_check.py 2>&1 | tail -15 | tee gpurun_out/multi8.log ;;
    bench_g8)
      for n in 2 4 8; do
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
          bench.py --gpus $n --steps 40 --warmup 5 2>gpurun_out/bench_g$n.err | tee gpurun_out/bench_g$n.json
      done ;;
    ncu_c5_full)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:mma_topk -s 4 -c 1 \
        -o gpurun_out/prof_c5 -f python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5_full.log 2>&1 ;;
    ncu_c5_list)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tav|mma_topk|threshold|finalize|query_prep|scan_rows|select|merge" -c 60 --csv \
        --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5_list.log 2>&1 ;;
    sanitizer)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
        -k "random_shapes or subset_duplicates or multi_pass or shard_merge or incremental" 2>&1 | tail -25 | tee gpurun_out/sanitizer_scan.log
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mma.py -m gpu -x -q \
        -k "search_matches_oracle or in_register or fallback" 2>&1 | tail -25 | tee gpurun_out/sanitizer_mma.log
      timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
        -k "batched_equals or multi_pass" 2>&1 | tail -15 | tee gpurun_out/sanitizer_race.log ;;
    bench_g8only)
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus 8 --steps 40 --warmup 5 2>gpurun_out/bench_g8.err | tee gpurun_out/bench_g8.json ;;
    bench_c4_g8)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 2>gpurun_out/bench_c4_g8.err | tee gpurun_out/bench_c4_g8.json ;;
    shard8)
      timeout 600 python bench.py --rows 1250000 --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_shard8.err | tee gpurun_out/bench_shard8.json
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:mma_topk -s 3 -c 1 \
        -o gpurun_out/prof_shard8 -f python bench.py --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_shard8.log 2>&1 ;;
    refbench)
      timeout 600 python tools/benchmark_vectorbase_gpu.py --json gpurun_out/benchmark_vectorbase_gpu.json 2>&1 | tail -40 | tee gpurun_out/benchmark_vectorbase_gpu.log ;;
    latency)
      timeout 600 python tools/latency_probe.py 2>&1 | tail -8 | tee gpurun_out/latency_probe.log ;;
    hypo)
      timeout 1200 python -m pytest tests/test_gpu_hypothesis.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_hypo.log ;;
    *) echo "unknown stage $st" ;;
  esac
done
echo "=== done ($(date +%T))"
