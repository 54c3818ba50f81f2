"""Split the SASS of one captured kernel at its BAR.SYNC instructions and print, per segment, the share of executed instructions
and of warp-stall samples (ncu --set full --import-source on; input: `ncu -i rep --page source --csv`).  usage: ncu_phase_stalls.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_")]
data = [(r[isrc].strip(), int(r[ia]), int(r[ist]), r) for r in rows[h + 1:] if len(r) > ia and r[ia].isdigit()]
tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
seg, n, st, first, reasons = 0, 0, 0, 0, Counter()
print(f"{len(data)} SASS instructions, {tot} executed, {tots} samples")
for i, (s, a, b, r) in enumerate(data):
    n += a
    st += b
    for ci, name in stall_cols:
        if r[ci].isdigit():
            reasons[name] += int(r[ci])
    if "BAR.SYNC" in s or "EXIT" in s and i == len(data) - 1 or i == len(data) - 1:
        top = ", ".join(f"{k[6:]} {v}" for k, v in reasons.most_common(3))
        print(f"segment {seg}: SASS {first}-{i}  inst {100.0 * n / tot:5.1f}%  samples {100.0 * st / tots:5.1f}%   [{top}]")
        seg, n, st, first, reasons = seg + 1, 0, 0, i + 1, Counter()
