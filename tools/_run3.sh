set -u
OUT=gpurun_out/${TAG:-r4b}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_parity.log
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new (levels 4)"
MMD_LEVELS=0 run "new levels 0"
MMD_LEVELS=2 run "new levels 2"
CASES="bf16_k100 fp8_k100 c3_k18"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library again"
run "new again"
grep sweep $OUT/sweep.log
for ph in 1 4; do
  for lv in 4 0; do
    MMD_LEVELS=$lv timeout 300 python bench.py --workload c5 --steps 4 --warmup 3 --phases $ph --no-cpu-baseline --no-extra > $OUT/bench_c5_ph${ph}_lv${lv}.json 2> $OUT/bench_c5_ph${ph}_lv${lv}.err
    python - <<PY
import json
try:
    p=json.loads(open("$OUT/bench_c5_ph${ph}_lv${lv}.json").read().strip().splitlines()[-1])
    print("c5 share phases=$ph levels=$lv:", round(p["ms_per_step"],1), "ms/step", round(p["roofline"]["achieved"]), "TFLOP/s", p["parity"]["recall_at_k"], p["clocks"]["sm_mhz"])
except Exception as e: print("c5 phases=$ph levels=$lv failed", e)
PY
  done
done
