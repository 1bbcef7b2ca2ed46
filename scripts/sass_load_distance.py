#!/usr/bin/env python
"""How far ahead of their first consumer are the hot loop's Tensor-Memory / shared-memory loads issued? (dev tool, no GPU)

    python scripts/sass_load_distance.py [libldrones.so] [-v] [--fn <substring of the mangled kernel name>] [--min-fp64 N]

Dumps the plain fp64 episode kernel's SASS (cuobjdump), finds the hot loop (the shortest backward branch that spans
>= 500 fp64-pipe instructions), decodes the scheduling control bits (stall count, write scoreboard, wait mask) and
prints, for every LDTM (tcgen05.ld) and LDS of the loop, the number of instructions and of statically scheduled issue
cycles between the load and the first instruction that waits on the load's scoreboard.  A load whose real latency is
longer than that distance stalls the warp (one warp per scheduler: nothing else can issue); ncu's per-instruction
`short_scoreboard` samples of the same build say by how much (profiles/*_stalls.txt).
"""
import re
import subprocess
import sys

args = sys.argv[1:]


def option(name, default):
    if name in args:
        i = args.index(name)
        v = args[i + 1]
        del args[i:i + 2]
        return v
    return default


PAT = option("--fn", "fly_mlp_f64_kernelILb0ELi0ELb0E")  # substring of the mangled kernel name
MIN_FP64 = int(option("--min-fp64", "500"))               # the hot loop spans at least this many fp64-pipe instructions
verbose = "-v" in args
lib = next((a for a in args if not a.startswith("-")), "learning_drones_b200/libldrones.so")
listing = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = next(ln.split("Function : ")[1].strip() for ln in listing.splitlines() if "Function" in ln and PAT in ln)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout

ins, cur = [], None
for ln in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
        ins.append(cur)
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)


def opcode(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
best = None
for a, text, _, _ in ins:
    if opcode(text) == "BRA":
        tgt = [x for x in text.split() if x.startswith("0x")]
        if tgt and int(tgt[-1], 16) < a:
            lo = int(tgt[-1], 16)
            n64 = sum(1 for b, t, _, _ in ins if lo <= b <= a and opcode(t) in FP64)
            if n64 >= MIN_FP64 and (best is None or a - lo < best[1] - best[0]):
                best = (lo, a)
win = [i for i in ins if best[0] <= i[0] <= best[1]]
stall = [(w1 >> 41) & 0xF for _, _, _, w1 in win]
wb = [(w1 >> 46) & 7 for _, _, _, w1 in win]
wait = [(w1 >> 52) & 0x3F for _, _, _, w1 in win]
n = len(win)
print(f"{lib}: hot loop {best[0]:#x}..{best[1]:#x}, {n} instructions, {sum(1 for i in win if opcode(i[1]) in FP64)} on the fp64 pipe, "
      f"static issue cycles {sum(stall)}")
tot = {}
for k, (a, text, _, _) in enumerate(win):
    op = opcode(text)
    if op not in ("LDTM", "LDS") or wb[k] == 7:
        continue
    d = cyc = 0
    j = k
    found = False
    for step in range(1, n + 1):  # wraps around the loop edge
        cyc += stall[j]
        j = (k + step) % n
        d += 1
        if wait[j] >> wb[k] & 1:
            found = True
            break
    tot.setdefault(op, []).append((d, cyc))
    if verbose or op == "LDTM":
        print(f"  [{k:4d}] {text:42s} sb{wb[k]} -> first wait after {d:3d} instructions / {cyc:3d} static cycles"
              f"{'' if found else ' (none)'}   at [{j:4d}] {win[j][1]}")
for op, v in tot.items():
    print(f"{op}: {len(v)} loads, distance to first wait: mean {sum(c for _, c in v) / len(v):.1f} static cycles, min {min(c for _, c in v)}, "
          f"below 40 cycles: {sum(1 for _, c in v if c < 40)}")
