#!/usr/bin/env python
"""Decode the scheduling control bits of a SASS listing (cuobjdump -sass) for an address range: per instruction the static
stall count, yield, scoreboard set / wait masks; prints the sum of the stall counts = the issue-to-issue cycles a single
warp needs for the range when no scoreboard wait bites (Volta-style 128-bit encoding, control field in bits 105..125).
    python scripts/sass_stalls.py kernel.sass 0x3d60 0x4940 [-v]"""
import re, sys
path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
verbose = "-v" in sys.argv
lines = open(path).read().splitlines()
out, i = [], 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            addr, text, w0, w1 = int(m.group(1), 16), m.group(2), int(m.group(3), 16), int(m2.group(1), 16)
            ctrl = (w1 >> 41) & 0x1fffff          # bits 105..125 of the 128-bit word
            stall, yld = ctrl & 0xf, (ctrl >> 4) & 1
            wr, rd, wait = (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f
            out.append((addr, text, stall, yld, wr, rd, wait))
            i += 2
            continue
    i += 1
sel = [o for o in out if lo <= o[0] <= hi]
tot = sum(o[2] for o in sel)
fp = sum(1 for o in sel if re.search(r'\bD(FMA|MUL|ADD|SETP)\b', o[1]))
waits = sum(1 for o in sel if o[6])
print(f"{len(sel)} instructions, {fp} FP64, sum of stall counts {tot} cycles ({tot/len(sel):.2f}/instr), {waits} with a scoreboard wait")
if verbose:
    for a, t, s, y, wr, rd, w in sel:
        print(f"{a:#07x} st{s:2d} {'Y' if y else ' '} wr{wr if wr != 7 else '-'} rd{rd if rd != 7 else '-'} wt{w:02x}  {t}")
