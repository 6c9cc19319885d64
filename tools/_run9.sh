set -u
OUT=gpurun_out/${TAG:-r4l}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
tr() { name=$1; shift; MMD_LIB_PATH=$D/libmmd_stats.so timeout 300 python tools/trace_run.py "$@" > $OUT/trace_$name.log 2>&1; echo "== $name ($*) ${MMD_LEVELS:-}"; grep "\[stats\]" $OUT/trace_$name.log; }
tr c3_n8share 16384 125000 768 18 text bf16
MMD_LEVELS=0 tr c3_n8share_lv0 16384 125000 768 18 text bf16
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
CASES="c3_n8share c3_k18"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library"
run "new"
MMD_LEVELS=0 run "new levels 0"
MMD_RESTART_TILES=8 run "new restart 8 tiles"
MMD_RESTART_TILES=16 run "new restart 16 tiles"
MMD_LIB_PATH=$D/libmmd_r1.so MMD_LIB_PARTIAL=1 run "r1 library again"
run "new again"
grep sweep $OUT/sweep.log
