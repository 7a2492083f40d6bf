#!/usr/bin/env python
"""Instruction mix from `ncu --page source --csv` output: executed warp instructions per SASS opcode (+ top source lines)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = 0
byop, samples = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        e, w = int(r[iE]), int(r[iW])
    except Exception:
        continue
    toks = r[iS].strip().split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    byop[op] += e
    samples[op] += w
    tot += e
print("total warp instructions", tot)
for op, e in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} {e:12d} {100 * e / tot:5.1f}%  stall samples {samples[op]}")
