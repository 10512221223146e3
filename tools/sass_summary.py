"""Per-kernel SASS summary of libyolo_b200.so: registers, spills, and the instruction mnemonics that prove the tcgen05 / TMEM /
TMA path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR / SYNCS = mbarrier traffic,
UTCATOMSWS = TMEM allocation).  Runs without a GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yolo_v3_b200", "lib", "libyolo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(v) for v in m.groups())
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "STL", "LDL", "R2UR"]
print(f"{'kernel':92s} {'instr':>6s} {'regs':>4s} {'stack':>5s} " + " ".join(f"{k:>7s}" for k in KEYS))
cnt = None
name = None
rows = []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cnt is not None:
        cnt["_n"] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
if name:
    rows.append((name, cnt))
for name, cnt in sorted(rows, key=lambda r: demangle(r[0])):
    d = demangle(name)
    d = re.sub(r"yb::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d)
    d = re.sub(r"^void ", "", d)
    u = usage.get(name, (0, 0, 0, 0))
    print(f"{d[:92]:92s} {cnt['_n']:6d} {u[0]:4d} {u[1]:5d} " + " ".join(f"{cnt[k]:7d}" for k in KEYS))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"\n{'total':92s} {tot['_n']:6d} {'':4s} {'':5s} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
