#!/usr/bin/env python
"""Summarise one gpurun profiling pass into profiles/ (tracked): launch list shares, the headline
ncu counters of the step kernel, the instruction mix and the stall breakdown.

    python scripts/summarize_ncu.py <tag>      # reads gpurun_out/{launches,prof,bench}_<tag>.*
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
out = []


def launches():
    p = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][-60:]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append(f"## Launch list (ncu --metrics gpu__time_duration.sum, cold-cache/serialised; shares matter)\n")
    out.append("| launches | total ms | share | kernel |\n|---:|---:|---:|---|")
    for k, a in agg.items():
        out.append(f"| {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.2f}% | `{k}` |")
    out.append("")


def raw():
    rep = os.path.join(G, f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return None
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    r = list(csv.reader(txt.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
    out.append("## Step kernel, ncu --set full (one launch)\n")
    out.append("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in d:
            out.append(f"| {w} | {d[w][0]} | {d[w][1]} |")
    stalls = sorted(((float(v[0]), h) for h, v in d.items()
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")), reverse=True)
    out.append("\nWarp stall reasons (warps stalled per issue-active cycle): " +
               ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in stalls[:8]) + "\n")
    def num(k):
        try:
            return float(d[k][0].replace(",", ""))
        except (KeyError, ValueError):
            return None
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd is not None and wr is not None:
        tb = rd * scale.get(d["dram__bytes_read.sum"][1], 1) + wr * scale.get(d["dram__bytes_write.sum"][1], 1)
        envs = None
        try:    # env count of the captured launch = the bench line of the same tag
            envs = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])["config"]["envs_per_gpu"]
        except (OSError, ValueError, KeyError, IndexError):
            pass
        json.dump({"tag": tag, "envs": envs, "dram_bytes_per_launch": tb, "kernel": d.get("Kernel Name", ("", ""))[0][:80]},
                  open(os.path.join(ROOT, "profiles", "traffic_opnav.json" if tag.startswith("opnav") else "traffic.json"), "w"))
        out.append(f"DRAM traffic per launch: {tb / 1e6:.1f} MB (read {rd} + write {wr} {d['dram__bytes_read.sum'][1]})\n")
    return rep


def source(rep):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = collections.Counter(); st = collections.Counter(); n = 0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in data:
        if len(r) < len(hdr):
            continue
        m = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
        ex[m[0].split(".")[0] if m else "?"] += int(r[ix["Instructions Executed"]] or 0); n += 1
        for s in stalls:
            st[s] += int(r[ix[s]] or 0)
    T = sum(ex.values()); S = sum(st.values()) or 1
    out.append(f"## SASS instruction mix (source page): {n} static instructions ({n * 16 / 1024:.0f} KiB), {T:.3e} executed warp instructions\n")
    out.append("| opcode | share of executed |\n|---|---:|")
    for k, v in ex.most_common(14):
        out.append(f"| {k} | {100 * v / T:.1f}% |")
    out.append("\nSampled stall reasons: " + ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in st.most_common(7)) + "\n")


out.append(f"# ncu summary {tag}\n")
for nm in (f"bench_{tag}.json",):
    p = os.path.join(G, nm)
    if os.path.exists(p) and os.path.getsize(p):
        out.append("Bench line of the same build (NOT taken under the profiler):\n\n```json\n" + open(p).read().strip() + "\n```\n")
launches()
rep = raw()
if rep:
    source(rep)
open(os.path.join(ROOT, "profiles", f"ncu_{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
