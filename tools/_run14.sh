set -u
OUT=gpurun_out/${TAG:-r4r}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_parity.log
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
CASES="q1 q100"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new"
grep sweep $OUT/sweep.log
bash tools/gpu_multi_check.sh ${TAG:-r4r} 2 c5
