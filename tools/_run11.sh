set -u
OUT=gpurun_out/${TAG:-r4o}; mkdir -p $OUT
export MMD_NO_AUTOBUILD=1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log; grep -i "recall" $OUT/pytest_gpu.log | head
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_c5_n1.json 2> $OUT/bench_c5_n1.err; echo "bench c5 rc=$?"
python - <<PY
import json
for f in ("bench_n1","bench_c5_n1"):
    try:
        p=json.loads(open("$OUT/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(p["value"]), "q/s", round(p["ms_per_step"],3), "ms; e2e", round(p["e2e"]["value"]), "; fused", round(p["roofline"]["achieved"]), "TF frac", round(p["roofline"]["frac"],3), "sust", round(p["roofline"]["frac_of_sustained"],3), p["parity"]["violations"], p["parity"]["recall_at_k"], p["clocks"]["sm_mhz"], p["config"]["rescore"][-60:])
        for k,v in p.get("also",{}).items(): print("  ",k, json.dumps(v)[:500])
        print("   cpu", json.dumps(p.get("cpu_baseline"))[:300])
    except Exception as e: print(f, "failed", e)
PY
