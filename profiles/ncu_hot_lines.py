#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: ncu_hot_lines.py rep kernel_name_substring [top_n]"""
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                     capture_output=True, text=True).stdout
hdr, cur_file, lines = None, None, []
STALLS = ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_lg',
          'stall_mio', 'stall_membar', 'stall_math', 'stall_no_inst']
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and len(r) < 5 and ('.cu' in r[1] or r[1].endswith('.h')):
        cur_file = r[1]
        continue
    if 'Instructions Executed' in r and 'Line No' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            st = {k: int(d.get(k, '0') or 0) for k in STALLS}
            lines.append((cur_file, int(r[0]), int(d['Instructions Executed']), int(d['# Samples']), st))
        except ValueError:
            pass
tot_i = sum(l[2] for l in lines) or 1
tot_s = sum(l[3] for l in lines) or 1
print(f"{kern}: total warp instructions {tot_i:,}  stall samples {tot_s:,}")
cache = {}
for f, ln, ins, smp, st in sorted(lines, key=lambda x: -(x[2] / tot_i + x[3] / tot_s))[:top]:
    path = f or ''
    local = path if os.path.exists(path) else os.path.join(ROOT, 'paimon_b200', 'csrc', os.path.basename(path))
    if local not in cache:
        try:
            cache[local] = open(local).read().splitlines()
        except Exception:
            cache[local] = []
    text = cache[local][ln - 1].strip()[:80] if 0 < ln <= len(cache[local]) else ''
    top_st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    sts = ' '.join(f"{k[6:]}={v}" for k, v in top_st if v)
    print(f"{os.path.basename(path):17s}:{ln:4d} inst {100 * ins / tot_i:5.1f}% stall {100 * smp / tot_s:5.1f}% [{sts}] {text}")
