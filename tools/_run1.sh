set -u
OUT=gpurun_out/r3c; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_n1.err
python tools/gpu_diag.py perf2 > $OUT/diag_perf2.log 2>&1; grep perf2 $OUT/diag_perf2.log
D=multimodal-misinformation-detection_b200/mmd_retrieval/dev
for W in w64 w32; do
  MMD_LIB_PATH=$D/libmmd_stats_$W.so python tools/trace_run.py 4096 1000000 768 18 text fp8 filter > $OUT/trace_filter_fp8_$W.log 2>&1
done
MMD_LIB_PATH=$D/libmmd_stats_w64.so python tools/trace_run.py 4096 1000000 768 18 text fp8 topk > $OUT/trace_k18_fp8.log 2>&1
MMD_LIB_PATH=$D/libmmd_stats_w64.so python tools/trace_run.py 4096 1000000 768 104 text fp8 topk > $OUT/trace_k104_fp8.log 2>&1
MMD_LIB_PATH=$D/libmmd_stats_w64.so python tools/trace_run.py 4096 1000000 768 104 text bf16 topk > $OUT/trace_k104_bf16.log 2>&1
MMD_LIB_PATH=$D/libmmd_stats_w64.so python tools/trace_run.py 4096 1000000 768 18 text bf16 topk > $OUT/trace_k18_bf16.log 2>&1
ls $OUT
