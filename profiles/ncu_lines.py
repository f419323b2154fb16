#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python ncu_lines.py x.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] == "":  continue            # SASS row; the line row above already aggregates
    d = dict(zip(hdr[4:], r[4:]))
    try:
        lines.append((cur_file, int(r[0]), r[1].strip()[:100], float(d["Instructions Executed"]), float(d["# Samples"]), d))
    except ValueError:
        pass
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
stall_keys = [k for k in hdr[4:] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: 0.0 for k in stall_keys}
for l in lines:
    for k in stall_keys:
        try: agg[k] += float(l[5][k])
        except Exception: pass
print("stalls:", ", ".join(f"{k[6:]}={v / max(ts,1) * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * ts))
for f, ln, src, i, s, d in lines:
    if i > thr / 100 * ti or s > thr / 100 * ts:
        top = sorted(((float(d[k]), k[6:]) for k in stall_keys if d.get(k, "0") not in ("", "0")), reverse=True)[:2]
        print(f"{f}:{ln:<4d} inst {i / ti * 100:5.1f}%  samp {s / ts * 100:5.1f}%  {' '.join(f'{k}:{v / max(s,1) * 100:.0f}%' for v, k in top):<28s} | {src}")
