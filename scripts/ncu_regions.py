#!/usr/bin/env python
"""Group the SASS lines of an `ncu --page source --csv` dump into runs of equal execution count.

    python scripts/ncu_regions.py gpurun_out/r2b_src.csv [ntiles]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntile = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kern = None; hdr = None; data = {}
for r in rows:
    if r and r[0] == 'Kernel Name': kern = r[1]; data[kern] = []; continue
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and kern and len(r) >= len(hdr) - 2: data[kern].append(r)
iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iSm = hdr.index('# Samples')
for k, d in data.items():
    tot = sum(int(r[iI]) for r in d); sm = sum(int(r[iSm]) for r in d)
    print(k, len(d), "inst", tot, "per tile", tot / ntile, "samples", sm)
    i = 0
    while i < len(d):
        j = i; c = int(d[i][iI])
        while j < len(d) and abs(int(d[j][iI]) - c) <= max(1, 0.02 * c): j += 1
        s = sum(int(r[iI]) for r in d[i:j]); smp = sum(int(r[iSm]) for r in d[i:j]); t = sum(int(r[iT]) for r in d[i:j])
        if s / tot > 0.004 or smp / sm > 0.01:
            print(f"{i:5d}-{j:5d} n={j-i:4d} exec/tile={c/ntile:8.1f} inst {s/tot*100:6.2f}% samples {smp/sm*100:6.2f}% thr {t/max(s,1):5.1f}  {d[i][iS][:50]}")
        i = j
