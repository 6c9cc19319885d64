set -u
export MMD_NO_AUTOBUILD=1
OUT=gpurun_out/r2_n8probe; mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/n8_probe.py > $OUT/probe_levels_default.log 2>&1; grep n8probe $OUT/probe_levels_default.log
MMD_LEVELS=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 tools/n8_probe.py > $OUT/probe_levels_0.log 2>&1; grep n8probe $OUT/probe_levels_0.log
