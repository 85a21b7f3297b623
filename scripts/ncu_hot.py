#!/usr/bin/env python
"""Top stall instructions of an `ncu --page source --csv` export (SASS view).
usage: python scripts/ncu_hot.py src.csv [N] [--range lo hi]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(body))
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:N]
for i in sorted(ranked):
    r = body[i]
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
    wf = r[ix['L1 Wavefronts Shared']] if 'L1 Wavefronts Shared' in ix else '-'
    ideal = r[ix['L1 Wavefronts Shared Ideal']] if 'L1 Wavefronts Shared Ideal' in ix else '-'
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}% ex={r[ix['Instructions Executed']]:>8} wf={wf:>9} ideal={ideal:>9} {r[ix['Source']].strip()[:60]:60s} {st}")
