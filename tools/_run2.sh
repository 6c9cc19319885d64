set -u
OUT=gpurun_out/r3e; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_parity.log
run() { SWEEP_TAG="$1" python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new w4 j2 (default)"
MMD_JOINT_AFTER=0 run "new w4 j0"
MMD_WINDOW=8 run "new w8 j2"
MMD_WINDOW=2 run "new w2 j2"
MMD_WINDOW=0 run "new w0 j2"
MMD_TRIGGER=16 run "new w4 j2 t16"
MMD_JOINT_AFTER=1 run "new w4 j1"
MMD_JOINT_AFTER=3 run "new w4 j3"
CASES="bf16_k100 fp8_k100 fp8_k104_4m c3_k18"
MMD_RESTART_TILES=16 run "new w4 j2 restart16"
MMD_RESTART_TILES=64 run "new w4 j2 restart64"
MMD_RESTART_TILES=200 run "new w4 j2 restart200"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library again"
grep sweep $OUT/sweep.log
