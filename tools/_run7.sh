set -u
N=${N:-4}
OUT=gpurun_out/${TAG:-r4j}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tools/e2e_probe.py > $OUT/probe_n$N.log 2>&1; echo "probe rc=$?"; grep "probe world" $OUT/probe_n$N.log
bash tools/gpu_multi_check.sh ${TAG:-r4j} $N
