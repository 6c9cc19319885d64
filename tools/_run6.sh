set -u
OUT=gpurun_out/${TAG:-r4i}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tools/e2e_probe.py > $OUT/probe_n$N.log 2>&1; echo "probe rc=$?"; grep "probe world" $OUT/probe_n$N.log
run() { SWEEP_TAG="$1" timeout 600 python tools/epi_sweep.py ${CASES:-} >> $OUT/sweep.log 2>&1; }
CASES="bf16_k100 fp8_k100 c3_k18"
MMD_EARLY=4 run "early 4"
run "early off (default)"
MMD_EARLY=4 run "early 4 again"
run "early off again"
grep sweep $OUT/sweep.log
