#!/usr/bin/env bash
# One gpurun call that produces the round's evidence under gpurun_out/<tag>/:
#   GPU parity tests, smoke(), the bench line (ours + reference arm), the ncu launch list and the
#   `--set full` captures of the fused kernel (C3 and C2 shapes) and of K1.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_evidence.sh <tag> [quick]
set -u
TAG=${1:-r1}
QUICK=${2:-}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > "$OUT/gpu.csv" 2>&1

export MMD_NO_AUTOBUILD=1
python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
# the pruning bounds are shared between CTAs as they are learnt (timing dependent): the parity file three more times
for rep in 1 2; do
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q > "$OUT/pytest_parity_rep$rep.log" 2>&1
  echo "parity repeat $rep rc=$?" | tee -a "$OUT/pytest_gpu.log"
done
python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1
echo "smoke rc=$?" | tee -a "$OUT/smoke.log"

python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
echo "bench rc=$?"
tail -c 600 "$OUT/bench_n1.json"
if [ -z "$QUICK" ]; then
  python bench.py --impl reference --steps 1 --warmup 0 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  echo "ref rc=$?"
  # the other BASELINE configs (c4 / c5: one GPU's 1/8 share of the 8-GPU configuration)
  for wl in c1 c2 c4 c5; do
    ST=30; [ $wl = c5 ] && ST=8
    python bench.py --workload $wl --steps $ST --warmup 5 --no-extra > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"
    echo "bench $wl rc=$?"
  done
  [ -n "${WITH_DIAG:-}" ] && python tools/gpu_diag.py perf2 k1perf > "$OUT/diag_perf.log" 2>&1
  D=${MMD_DEV_LIBS:-multimodal-misinformation-detection_b200/mmd_retrieval/dev}   # developer builds (tools/build_stats.sh; a round-1 libmmd_r1.so for A/B sweeps), shipped to the box only when placed there
  SWEEP_TAG="this library" python tools/epi_sweep.py > "$OUT/sweep.log" 2>&1
  [ -f $D/libmmd_r1.so ] && MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 SWEEP_TAG="round-1 library" python tools/epi_sweep.py >> "$OUT/sweep.log" 2>&1
  grep sweep "$OUT/sweep.log"
  if [ -f $D/libmmd_stats.so ] && [ -n "${WITH_COUNTERS:-}" ]; then
    for sh in "c3_k18 16384 1000000 768 18 text bf16" "fp8_k18 16384 1000000 768 18 text fp8" "bf16_k104 16384 1000000 768 104 text bf16" \
              "fp8_k104 16384 1000000 768 104 text fp8" "c2_k18 4096 50000 2048 18 image bf16" "c3_n8share 16384 125000 768 18 text bf16"; do
      set -- $sh; name=$1; shift
      MMD_LIB_PATH=$D/libmmd_stats.so python tools/trace_run.py "$@" > "$OUT/trace_$name.log" 2>&1
      echo "== $name" >> "$OUT/phase_counters.log"; grep "\[stats\]" "$OUT/trace_$name.log" >> "$OUT/phase_counters.log"
    done
    cat "$OUT/phase_counters.log"
  fi
fi

SHORT="--steps 2 --warmup 1 --no-cpu-baseline --no-extra"
# launch list of one short run (shares, not absolutes)
python bench.py $SHORT > "$OUT/plain_c3.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_c3.csv" \
    python bench.py $SHORT > "$OUT/ncu_launches.log" 2>&1
# full capture of the dominant kernel, C3 shape
ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -s 1 -c 2 -f -o "$OUT/prof_c3_fused" \
    python bench.py $SHORT > "$OUT/ncu_c3_fused.log" 2>&1
if [ -z "$QUICK" ]; then
  # C2 shape
  ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -s 1 -c 2 -f -o "$OUT/prof_c2_fused" \
      python bench.py $SHORT --workload c2 > "$OUT/ncu_c2_fused.log" 2>&1
  # fp8 instantiations of the fused kernel: CAP = 64 (K' = 18) and CAP = 128 (K' = 100), CTA pairs
  for cs in fp8_k18 fp8_k100; do
    ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -s 3 -c 1 -f -o "$OUT/prof_${cs}_fused" \
        python tools/epi_sweep.py $cs > "$OUT/ncu_${cs}_fused.log" 2>&1
  done
  # K1 (corpus prepare is the first normalize_cast launch) + rescore + merge
  ncu --set full --clock-control none -k regex:"normalize_cast|rescore|merge" -c 5 -f -o "$OUT/prof_c3_aux" \
      python bench.py $SHORT > "$OUT/ncu_c3_aux.log" 2>&1
fi
# gpurun brings back at most 64 MiB: summarise the captures here, keep only the C3 capture itself (source page, read offline)
python tools/make_profiles.py "$OUT" r2 "$OUT/summ" > "$OUT/make_profiles.log" 2>&1
rm -f "$OUT"/prof_c2_fused.ncu-rep "$OUT"/prof_c3_aux.ncu-rep "$OUT"/prof_fp8_*.ncu-rep
ls -la "$OUT" "$OUT/summ"
