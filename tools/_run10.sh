set -u
OUT=gpurun_out/${TAG:-r4m}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_parity.log
tr() { name=$1; shift; MMD_LIB_PATH=$D/libmmd_stats.so timeout 300 python tools/trace_run.py "$@" > $OUT/trace_$name.log 2>&1; echo "== $name ($*) ${MMD_LEVELS:-}"; grep "\[stats\]" $OUT/trace_$name.log; }
tr c3_n8share 16384 125000 768 18 text bf16
tr c3_k18 16384 1000000 768 18 text bf16
tr bf16_k104 16384 1000000 768 104 text bf16
tr fp8_k104 16384 1000000 768 104 text fp8
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new"
CASES="c3_n8share c3_k18 bf16_k100 fp8_k100 fp8_k18 c2_k18"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library again"
run "new again"
grep sweep $OUT/sweep.log
