"""Executed-instruction share per basic block from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
data = [(r[0], r[isrc].strip(), int(r[ia]), int(r[ismp])) for r in rows[2:] if len(r) > ia]
tot = sum(d[2] for d in data); smptot = sum(d[3] for d in data)
print("kernel:", rows[0][1]); print("total warp-instructions", tot, "sass lines", len(data), "samples", smptot)
out = []; cur = None
for i, (a, s, c, sm) in enumerate(data):
    if cur and cur[2] == c: cur[1] = i; cur[3] += sm
    else: cur = [i, i, c, sm]; out.append(cur)
for a, b, c, sm in out:
    n = b - a + 1
    if c * n / tot > thr:
        print(f"{a:5d}-{b:5d} n={n:4d} exec={c:>14,d} inst-share={100*c*n/tot:5.1f}%  sample-share={100*sm/max(smptot,1):5.1f}%   {data[a][1][:50]}")
