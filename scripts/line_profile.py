#!/usr/bin/env python
"""Join an ncu source-page CSV (per-SASS-instruction samples / executed counts) with nvdisasm's inline line
info of the same cubin and aggregate per source line of leo_core.cuh.

    nvcc ... -cubin -o /tmp/bskenv.cubin csrc/bskenv.cu            (scripts/sass_loops.sh does this)
    ncu -i prof.ncu-rep --page source --csv > src.csv
    python scripts/line_profile.py /tmp/bskenv.cubin src.csv [kernel-substring] [depth]

depth 0 = innermost frame (the line the instruction was generated from), 1 = its caller, ...; -1 = outermost
frame inside leo_core.cuh (the line of the top-level function that led to the instruction)."""
import collections
import csv
import re
import subprocess
import sys

cubin, src = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "leo_step_kernelILi3ELb0ELb1"
depth = int(sys.argv[4]) if len(sys.argv) > 4 else -1
import os
W, T = float(os.environ.get("WARPS", 4096.0)), float(os.environ.get("TICKS", 1800.0))
CORE = os.environ.get("CORE", "leo_core.cuh")     # CORE=opnav_core.cuh WARPS=1024 TICKS=3000 for the opNav kernel
only = os.environ.get("SECTION")     # restrict the per-line table to one out-of-line function

txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
chains, cur, on = {}, [], False
sub, subs = "kernel", {}
for line in txt.splitlines():
    if line.startswith(".text."):
        on = kern in line
        continue
    if not on:
        continue
    m = re.match(r"^\$\S*\$(\S+):\s*$", line)          # out-of-line device function / libdevice slow path
    if m:
        nm = m.group(1)
        mm = re.search(r"(?:3leo|5opnav)(\d+)", nm)
        sub = nm[mm.end():mm.end() + int(mm.group(1))] if mm else nm.strip("_$")[:40]
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur.append((m.group(1).rsplit("/", 1)[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        a = int(m.group(1), 16)
        if cur:
            chains["last"] = cur
        chains[a] = chains.get("last", [])
        subs[a] = sub
        cur = []

rows = list(csv.reader(open(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
sagg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16)
    if base is None:
        base = a
    ch = chains.get(a - base, [])
    core = [c for c in ch if c[0] == CORE]
    if depth == -1:
        key = core[-1] if core else (ch[-1] if ch else ("?", 0))
    else:
        key = ch[min(depth, len(ch) - 1)] if ch else ("?", 0)
    n = int(r[ix["# Samples"]] or 0)
    ex = int(r[ix["Instructions Executed"]] or 0) / W / T
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    sname = subs.get(a - base, "?")
    if only and sname != only:
        tot += int(r[ix["# Samples"]] or 0)
        continue
    sagg[sname][0] += n; sagg[sname][1] += ex
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        sagg[sname][2] += ex
    agg[key][0] += n
    agg[key][1] += ex
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        agg[key][2] += ex
    tot += n
lines = {}
try:
    lines = dict(enumerate(open("/root/repo/basilisk_env_b200/csrc/" + CORE).read().splitlines(), 1))
except OSError:
    pass
# per-function totals (function = the last definition that starts at or before the line)
fstarts = []
for ln, text in sorted(lines.items()):
    m = re.match(r"^(?:LEO_HD_NOINLINE|LEO_HD|ON_HD_NOINLINE|ON_HD)\s+.*?(\w+)\(", text)
    if m and not text.startswith(" "):
        fstarts.append((ln, m.group(1)))
def func_of(key):
    if key[0] != CORE:
        return key[0]
    name = "?"
    for ln, n in fstarts:
        if ln <= key[1]:
            name = n
    return name
fagg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for key, (n, ex, fp) in agg.items():
    f = func_of(key)
    fagg[f][0] += n; fagg[f][1] += ex; fagg[f][2] += fp
print(f"{'code section':>28s} {'time%':>6s} {'instr/tick':>10s} {'fp64/tick':>9s}")
for f, (n, ex, fp) in sorted(sagg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{f:>28s} {100 * n / tot:6.2f} {ex:10.1f} {fp:9.1f}")
print()
print(f"{'function':>28s} {'time%':>6s} {'instr/tick':>10s} {'fp64/tick':>9s}")
for f, (n, ex, fp) in sorted(fagg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{f:>28s} {100 * n / tot:6.2f} {ex:10.1f} {fp:9.1f}")
print()
print(f"{'line':>16s} {'time%':>6s} {'instr/tick':>10s} {'fp64/tick':>9s}")
for key, (n, ex, fp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
    text = lines.get(key[1], "").strip()[:90] if key[0] == CORE else ""
    print(f"{key[0][:10]:>10s}:{key[1]:<5d} {100 * n / tot:6.2f} {ex:10.1f} {fp:9.1f}  {text}")
