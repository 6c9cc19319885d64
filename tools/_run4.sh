set -u
OUT=gpurun_out/${TAG:-r4c}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_parity.log
tr() { name=$1; shift; MMD_LIB_PATH=$D/libmmd_stats.so timeout 300 python tools/trace_run.py "$@" > $OUT/trace_$name.log 2>&1; echo "== $name ($*) ${MMD_LEVELS:-}"; grep "\[stats\]" $OUT/trace_$name.log; }
tr bf16_k104 16384 1000000 768 104 text bf16
MMD_LEVELS=0 tr bf16_k104_lv0 16384 1000000 768 104 text bf16
tr fp8_k104 16384 1000000 768 104 text fp8
tr c3_k18 16384 1000000 768 18 text bf16
tr fp8_k18 16384 1000000 768 18 text fp8
tr c2_k18 4096 50000 2048 18 image bf16
MMD_LEVELS=0 tr c2_k18_lv0 4096 50000 2048 18 image bf16
tr q100 100 1000000 768 18 text bf16
