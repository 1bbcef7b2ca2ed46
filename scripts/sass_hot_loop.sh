#!/bin/bash
# dev tool: dump the plain fp64 episode kernel's SASS from a built library and print the hot loop's opcode / issue-slot
# histogram (scripts/sass_hist.py).  usage: sass_hot_loop.sh [lib.so] [out.sass]
LIB=${1:-learning_drones_b200/libldrones.so}
OUT=${2:-/tmp/fly_plain.sass}
FN=$(cuobjdump -sass "$LIB" 2>/dev/null | grep "Function" | grep "fly_mlp_f64_kernelILb0ELi0ELb0E" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$FN" "$LIB" > "$OUT" 2>&1
RANGE=$(python3 - "$OUT" <<'PY'
import re, sys
ops, bras = [], []
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m:
        continue
    a, t = int(m.group(1), 16), m.group(2).split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops.append((a, op))
    if op == "BRA":
        tgt = [x for x in t if x.startswith("0x")]
        if tgt and int(tgt[-1], 16) < a:
            bras.append((int(tgt[-1], 16), a))
best = None
for lo, hi in bras:  # the hot loop = the shortest backward branch that spans >= 500 fp64-pipe instructions
    n64 = sum(1 for a, op in ops if lo <= a <= hi and op in ("DFMA", "DMUL", "DADD"))
    if n64 >= 500 and (best is None or hi - lo < best[1] - best[0]):
        best = (lo, hi)
print(hex(best[0]), hex(best[1] + 16))
PY
)
python3 "$(dirname "$0")/sass_hist.py" "$OUT" $RANGE
