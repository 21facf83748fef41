#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: ncu_hot_lines.py rep kernel_name_substring [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur_file, lines, src = None, None, [], {}
for r in rows:
    if len(r) >= 2 and r[0].strip() == 'File Name':
        cur_file = r[1]; continue
    if 'Instructions Executed' in r and 'Line No' in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            lines.append((cur_file, int(r[0]), int(d['Instructions Executed']), int(d['# Samples'])))
        except ValueError:
            pass
tot_i = sum(l[2] for l in lines) or 1
tot_s = sum(l[3] for l in lines) or 1
print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
files = {}
for f, ln, ins, smp in sorted(lines, key=lambda x: -x[2])[:top]:
    if f not in files:
        try: files[f] = open(f).read().splitlines()
        except Exception: files[f] = []
    text = files[f][ln - 1].strip()[:95] if ln - 1 < len(files[f]) else ''
    print(f"{(f or '?').split('/')[-1]:18s}:{ln:4d} inst {100*ins/tot_i:5.1f}%  stall {100*smp/tot_s:5.1f}%  {text}")
