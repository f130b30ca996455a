#!/usr/bin/env python
"""Quick look at one .ncu-rep: headline counters, instruction mix per tick, stall breakdown."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; warps = float(sys.argv[2]) if len(sys.argv) > 2 else 4096; ticks = 1800
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
r = list(csv.reader(txt.splitlines())); d = dict(zip(r[0], r[2]))
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    print(f"{k:75s} {d.get(k)}")
st = sorted(((float(v), k.split('stalled_')[1].split('_per_issue')[0]) for k, v in d.items()
             if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
print("stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(txt.splitlines())); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); n = 0
for row in rows[2:]:
    if len(row) < len(hdr): continue
    m = [o for o in row[ix["Source"]].split() if not o.startswith("@")]
    ex[m[0].split(".")[0] if m else "?"] += int(row[ix["Instructions Executed"]] or 0); n += 1
T = sum(ex.values())
print(f"static {n} instrs ({n*16/1024:.0f} KiB); executed per warp-tick {T/warps/ticks:.0f}")
print(", ".join(f"{k} {v/warps/ticks:.0f}" for k, v in ex.most_common(16)))
