#!/usr/bin/env python
"""Opcode histogram + issue-slot budget of a SASS address window (dev tool).

    cuobjdump -sass -fun <mangled kernel> libldrones.so > k.sass
    python scripts/sass_hist.py k.sass 0x6000 0x9e60

Decodes the scheduling control bits of every 128-bit instruction (sm_70+: high word bits 41-44 = stall count,
45 = yield, 52-57 = scoreboard wait mask) and prints per-opcode counts, the sum of the static stall counts (the
issue cycles ONE warp needs for the window when no scoreboard wait triggers), and the fp64-pipe share.
"""
import collections
import re
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
ins = []
cur = None
for ln in open(path):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
        ins.append(cur)
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)
win = [i for i in ins if lo <= i[0] < hi]
ops = collections.Counter()
stall_by_op = collections.Counter()
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
total_stall = fp64 = 0
follow = collections.Counter()
for k, (addr, text, w0, w1) in enumerate(win):
    t = text.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    stall = (w1 >> 41) & 0xF
    ops[op] += 1
    stall_by_op[op] += stall
    total_stall += stall
    if op in FP64:
        fp64 += 1
        nxt = win[k + 1][1].split() if k + 1 < len(win) else ["?"]
        nop = (nxt[1] if nxt[0].startswith("@") else nxt[0]).split(".")[0]
        follow[("fp64" if nop in FP64 else "other", stall)] += 1
print(f"window {lo:#x}..{hi:#x}: {len(win)} instructions, {fp64} on the fp64 pipe, sum of static stall counts {total_stall}")
for op, n in ops.most_common():
    print(f"  {op:12s} {n:5d}   mean stall {stall_by_op[op] / n:4.2f}")
print("stall count after an fp64 instruction, by what follows:", dict(sorted(follow.items())))
