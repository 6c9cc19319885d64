"""Opcode evidence for the built library: per kernel, how many tcgen05 / TMEM / TMA / mbarrier / legacy-tensor instructions
its SASS holds (the PTX names never appear in SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor ->
UTMALDG, mma.sync -> HMMA).  python tools/sass_ops.py > profiles/sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-misinformation-detection_b200", "mmd_retrieval", "libmmd.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "QGMMA", "HGMMA", "IMMA",
         "ATOM", "RED", "LDS", "STS", "SHFL", "VOTE", "FMNMX", "FMNMX3", "LDG", "STG", "MEMBAR", "ERRBAR"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:  # noqa: BLE001
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    info = open(LIB + ".buildinfo").read().strip() if os.path.exists(LIB + ".buildinfo") else "?"
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   [{info}]")
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
                    total[w] += 1
    print("# library totals: " + ", ".join(f"{w}={total[w]}" for w in WATCH if total[w]))
    for fn, c in counts.items():
        name = demangle(fn)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        ops = ", ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{name}\n    instructions={c['_all']}  {ops}")


if __name__ == "__main__":
    sys.exit(main())
