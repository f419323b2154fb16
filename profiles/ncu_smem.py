#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts from an `ncu --page source --print-source cuda,sass --csv` dump.
usage: python ncu_smem.py x.csv frames"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); frames = float(sys.argv[2])
hdr = None; cur = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        w = float(d["L1 Wavefronts Shared"]); ex = float(d["L1 Wavefronts Shared Excessive"]); ideal = float(d["L1 Wavefronts Shared Ideal"])
    except Exception: continue
    if w > 0: out.append((cur, int(r[0]), w, ex, ideal, r[1].strip()[:90]))
tot = sum(o[2] for o in out)
print("total shared wavefronts/frame", round(tot / frames, 1))
for o in sorted(out, key=lambda o: -o[2])[:24]:
    print(f"{o[0]}:{o[1]:<4d} wf/frame {o[2]/frames:7.1f} excessive {o[3]/frames:6.1f} ideal {o[4]/frames:7.1f} | {o[5]}")
