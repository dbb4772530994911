#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
data = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) // 2: continue
    if r[2] != "-": continue          # SASS rows carry an address; per-line summary rows have "-"
    try:
        ii = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
        data.append((int(r[ii]), int(r[sa]), cur_file, r[0], r[1].strip()[:100]))
    except Exception:
        pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print(f"total warp-instructions {tot:,}  samples {tots:,}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/tot*100:5.1f}% inst {d[1]/max(tots,1)*100:5.1f}% smp  {d[2]}:{d[3]:>4}: {d[4]}")
