set -u
OUT=gpurun_out/${TAG:-r4p}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
CASES="q1 q100 c2_k18"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new"
MMD_LEVELS=0 run "new levels 0"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library again"
run "new again"
MMD_LEVELS=0 run "new levels 0 again"
grep sweep $OUT/sweep.log
