set -u
N=8
OUT=gpurun_out/${TAG:-r4k}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tools/e2e_probe.py > $OUT/probe_n$N.log 2>&1; echo "probe rc=$?"; grep "probe world" $OUT/probe_n$N.log
timeout 200 python bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_c3_n$N.json 2> $OUT/bench_c3_n$N.err; echo "bench rc=$?"
python - <<PY
import json
p=json.loads(open("$OUT/bench_c3_n$N.json").read().strip().splitlines()[-1])
print(p['value'], p['ms_per_step'], p['e2e'], p['roofline']['kernel_ms'], p['parity'])
PY
