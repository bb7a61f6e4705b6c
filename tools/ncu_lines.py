#!/usr/bin/env python
"""Per-source-line totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`: instructions and stall samples."""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = r[2]
ci = hdr.index('Instructions Executed'); cs = hdr.index('# Samples')
lines = {}; cur = None; order = []
for row in r[3:]:
    if len(row) < 8 or row[0] == 'Line No': continue
    if row[0] != '':
        cur = (row[0], row[1]); 
        if cur not in lines: lines[cur] = [0, 0]; order.append(cur)
        try: pass
        except: pass
    try: lines[cur][0] += int(row[ci]); lines[cur][1] += int(row[cs])
    except Exception: pass
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print('total warp instructions', tot, 'samples', ts)
for k in order:
    v = lines[k]
    if v[0] > tot * thr / 100 or v[1] > ts * thr / 100:
        print(f"{k[0]:>5} {100*v[0]/tot:5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp | {k[1][:140]}")
