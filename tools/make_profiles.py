#!/usr/bin/env python
"""Turn one evidence run (gpurun_out/<tag>/, produced by tools/gpu_evidence.sh) into the tracked files under
profiles/: the launch list, one text summary per .ncu-rep, and traffic.json (dram bytes per fused launch) that
bench.py reads for roofline.traffic.

    python tools/make_profiles.py gpurun_out/r1a r1 [destination directory, default profiles/]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
]


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return float(val) * scale if scale else None


def summarise(path, out_md):
    hdr, units, launches = raw_page(path)
    traffic = []
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full --clock-control none : {os.path.basename(path)}\n\n")
        f.write("(read offline with `ncu -i <rep> --page raw --csv`; one block per profiled launch; per-launch times here "
                "are cold-cache, serialised replays)\n\n")
        for row in launches:
            d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?').split('(')[0][-90:]}  grid={d.get('Grid Size')} block={d.get('Block Size')}\n\n```\n")
            rd = wr = None
            for h in hdr:
                base = h.split(".TriageCompute.")[-1]
                if base in KEYS:
                    f.write(f"{base:82s} {d[h]:>18s} {u[h]}\n")
                if base == "dram__bytes_read.sum":
                    rd = to_bytes(d[h], u[h])
                if base == "dram__bytes_write.sum":
                    wr = to_bytes(d[h], u[h])
            f.write("```\n\n")
            traffic.append({"kernel": d.get("Kernel Name", "?").split("(")[0][-60:], "dram_read_bytes": rd, "dram_write_bytes": wr})
    return traffic


def launches_table(csv_path, out_md, title):
    rows = list(csv.reader(open(csv_path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[start]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    total = {}
    lines = []
    for r in rows[start + 1:]:
        name = r[ki].split("(")[0]
        short = name.split("::")[-1][-60:]
        ns = float(r[vi].replace(",", ""))
        lines.append((r[0], short, r[gi], r[bi], ns))
        total[short] = total.get(short, 0.0) + ns
    all_ns = sum(total.values())
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n## time share per kernel\n\n| kernel | launches' total (us) | share |\n|---|---|---|\n")
        for k, v in sorted(total.items(), key=lambda t: -t[1]):
            f.write(f"| {k} | {v / 1e3:.1f} | {100 * v / all_ns:.2f} % |\n")
        f.write("\n## every launch\n\n| id | kernel | grid | block | ns |\n|---|---|---|---|---|\n")
        for l in lines:
            f.write(f"| {l[0]} | {l[1]} | {l[2]} | {l[3]} | {l[4]:.0f} |\n")


def main():
    src, rnd = sys.argv[1], sys.argv[2]
    dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")     # (on the GPU box: a directory under gpurun_out/)
    os.makedirs(dst, exist_ok=True)
    traffic_json = {}
    tpath = os.path.join(dst, "traffic.json")
    if os.path.exists(tpath):
        traffic_json = json.load(open(tpath))
    for name in sorted(os.listdir(src)):
        p = os.path.join(src, name)
        if name.startswith("launches") and name.endswith(".csv"):
            launches_table(p, os.path.join(dst, f"{rnd}_{name[:-4]}.md"), f"{rnd}: launch list of `bench.py --steps 2 --warmup 1` ({name})")
        if name.endswith(".ncu-rep"):
            tr = summarise(p, os.path.join(dst, f"{rnd}_{name[:-8]}.md"))
            for wl in ("c3", "c2"):
                if name == f"prof_{wl}_fused.ncu-rep" and tr:
                    vals = [t["dram_read_bytes"] + t["dram_write_bytes"] for t in tr if t["dram_read_bytes"] is not None]
                    if vals:
                        traffic_json[wl] = sum(vals) / len(vals)
        if name in ("bench_n1.json", "bench_ref.json", "pytest_gpu.log", "smoke.log", "diag_perf.log") or \
                (name.startswith("bench_c") and name.endswith(".json")):
            with open(p) as fi, open(os.path.join(dst, f"{rnd}_{name}"), "w") as fo:
                fo.write(fi.read())
    json.dump(traffic_json, open(tpath, "w"), indent=1)
    print("traffic:", traffic_json)


if __name__ == "__main__":
    main()
