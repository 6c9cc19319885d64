set -u
OUT=gpurun_out/${TAG:-r4q}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
tr() { name=$1; shift; MMD_LIB_PATH=$D/libmmd_stats.so timeout 300 python tools/trace_run.py "$@" > $OUT/trace_$name.log 2>&1; echo "== $name ($*) ${MMD_LEVELS:-}"; grep "\[stats\]" $OUT/trace_$name.log; grep "\[trace\] tile" $OUT/trace_$name.log | sed -n '1,6p;20,26p'; }
tr q1 1 1000000 768 18 text bf16
tr q100 100 1000000 768 18 text bf16
