#!/usr/bin/env bash
# Multi-GPU evidence for one world size: the >= 2-GPU parity tests and the C3 bench line.
# Usage (repo root, GPU box with >= N GPUs): bash tools/gpu_multi_check.sh <tag> <N> [extra bench workloads...]
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
MMD_TEST_WORLD=$N MMD_EXPECT_EXCHANGE=peer timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > "$OUT/pytest_multi_n$N.log" 2>&1
echo "pytest multi (world $N) rc=$?" | tee -a "$OUT/pytest_multi_n$N.log"
tail -5 "$OUT/pytest_multi_n$N.log"
timeout 600 python bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench_c3_n$N.json" 2> "$OUT/bench_c3_n$N.err"
echo "bench c3 n$N rc=$?"; tail -c 1500 "$OUT/bench_c3_n$N.json"; tail -5 "$OUT/bench_c3_n$N.err"
for wl in "$@"; do
  timeout 900 python bench.py --gpus $N --workload $wl --steps 8 --warmup 3 > "$OUT/bench_${wl}_n$N.json" 2> "$OUT/bench_${wl}_n$N.err"
  echo "bench $wl n$N rc=$?"; tail -c 1500 "$OUT/bench_${wl}_n$N.json"; tail -5 "$OUT/bench_${wl}_n$N.err"
done
