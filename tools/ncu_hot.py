#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu --page source --csv` (SASS view): address, samples, instruction."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break  # first launch only
    body.append(r)
data = [(int(r[iS] or 0), i, r[iSrc].strip(), int(r[iEx] or 0)) for i, r in enumerate(body) if len(r) > iS and r[iS].isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for smp, i, src, ex in sorted(top, key=lambda t: t[1]):
    print("%5d %5.1f%%  line %5d  exec %9d  %s" % (smp, 100.0 * smp / tot, i, ex, src[:100]))
