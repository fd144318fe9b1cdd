#!/usr/bin/env python3
"""Hot SASS lines of one kernel from `ncu -i rep --page source --csv` output (stdin or file): share of stall samples,
executed warp instructions and shared-memory wavefronts per instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[h + 1:] if len(r) > 10 and r[0] != "Address"]
def I(r, n):
    try: return int(r[idx[n]])
    except Exception: return 0
ts = sum(I(r, '# Samples') for r in body); ti = sum(I(r, 'Instructions Executed') for r in body); tw = sum(I(r, 'L1 Wavefronts Shared') for r in body)
print(f"samples {ts}  warp-instructions {ti/1e6:.1f}M  smem wavefronts {tw/1e6:.1f}M")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
for k, r in enumerate(body):
    s, i, w, wi = I(r, '# Samples'), I(r, 'Instructions Executed'), I(r, 'L1 Wavefronts Shared'), I(r, 'L1 Wavefronts Shared Ideal')
    if s > ts * thr or (tw and w > tw * 0.02):
        top = sorted(((I(r, n), n[6:]) for n in stalls), reverse=True)[:2]
        print(f"{k:4d} {r[idx['Source']].strip()[:58]:58s} samp {100*s/max(ts,1):5.1f}% inst {i/1e6:7.2f}M wf {w/1e6:6.2f}M (ideal {wi/1e6:6.2f}M) {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
