# 8-GPU evidence at HEAD: the >= 2-GPU parity tests at world 8 (text and joint + fp8; the small-shard stress ran at worlds 2, 4
# and 8 earlier in the round), then the bench lines of C3 and of the two configs defined on 8 GPUs.
set -u
export MMD_NO_AUTOBUILD=1
TAG=r2_n8; N=8
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
MMD_TEST_WORLD=$N MMD_EXPECT_EXCHANGE=peer timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "multi_gpu or joint_and_fp8" > "$OUT/pytest_multi_n$N.log" 2>&1
echo "pytest multi (world $N) rc=$?" | tee -a "$OUT/pytest_multi_n$N.log"; tail -3 "$OUT/pytest_multi_n$N.log"
timeout 300 python bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench_c3_n$N.json" 2> "$OUT/bench_c3_n$N.err"; echo "bench c3 rc=$?"
timeout 400 python bench.py --gpus $N --workload c5 --steps 8 --warmup 3 > "$OUT/bench_c5_n$N.json" 2> "$OUT/bench_c5_n$N.err"; echo "bench c5 rc=$?"
timeout 300 python bench.py --gpus $N --workload c4 --steps 8 --warmup 3 > "$OUT/bench_c4_n$N.json" 2> "$OUT/bench_c4_n$N.err"; echo "bench c4 rc=$?"
python - <<PY
import json
for w in ("c3","c5","c4"):
    try:
        p=json.loads(open("$OUT/bench_%s_n8.json"%w).read().strip().splitlines()[-1])
        print(w, round(p["value"]), "q/s", round(p["ms_per_step"],3), "ms; e2e", round(p["e2e"]["value"]), round(p["e2e"]["ms_per_step"],3), "; kernel", round(p["roofline"]["kernel_ms"],3), round(p["roofline"]["achieved"]), "TF frac", round(p["roofline"]["frac"],3), "parity", p["parity"]["violations"], p["parity"]["recall_at_k"], p["clocks"]["sm_mhz"], p["config"]["exchange"])
    except Exception as e: print(w, "failed", e)
PY
