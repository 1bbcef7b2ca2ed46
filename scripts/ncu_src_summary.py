"""Summarise an `ncu --page source --csv` dump: per-opcode executed counts / stall samples, and
per-address-window stall breakdown (dev tool).  usage: ncu_src_summary.py file.csv [warp_steps]"""
import csv, collections, sys
f = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else 4.096e6
rows = list(csv.reader(open(f))); hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ops = collections.Counter(); samp = collections.Counter(); stall_tot = collections.Counter()
recs = []
for r in data:
    src = r[col['Source']].strip(); n = int(r[col['Instructions Executed']]); s = int(r[col['# Samples']])
    tt = src.split(); op = (tt[1] if tt[0].startswith('@') else tt[0]).split('.')[0]
    ops[op] += n; samp[op] += s
    st = {h: int(r[col[h]]) for h in stall_cols}
    for h, v in st.items(): stall_tot[h] += v
    recs.append((src, n, s, st))
tot = sum(ops.values()); ts = sum(samp.values())
print(f"instr per warp-step {tot/steps:.1f}; samples {ts}")
for op, n in ops.most_common(16):
    print(f"  {op:8s} {n/steps:7.1f}/step  samples {100*samp[op]/ts:5.1f}%")
print("stall totals:", {k[6:]: f"{100*v/ts:.1f}%" for k, v in stall_tot.most_common(8)})
# windows of 100 instructions over the hot region (executed >= 0.5*steps)
hot = [(i, r) for i, r in enumerate(recs) if r[1] >= 0.5 * steps]
W = 60
for w in range(0, len(hot), W):
    chunk = hot[w:w + W]
    s = sum(r[2] for _, r in chunk)
    st = collections.Counter()
    for _, r in chunk:
        for k, v in r[3].items(): st[k[6:]] += v
    kinds = collections.Counter((r[0].split()[1] if r[0].startswith('@') else r[0].split()[0]).split('.')[0] for _, r in chunk)
    top = ", ".join(f"{k}:{100*v/max(s,1):.0f}%" for k, v in st.most_common(3))
    print(f"  hot[{w:4d}:{w+len(chunk):4d}] samples {100*s/ts:5.1f}%  cyc/inst {s/ts*len(hot)/len(chunk):4.2f}x  {top}   {dict(kinds.most_common(4))}")
