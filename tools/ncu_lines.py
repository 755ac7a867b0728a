#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from a .ncu-rep taken with --import-source on
(build with -lineinfo):   python tools/ncu_lines.py gpurun_out/x.ncu-rep k_candidates [min_pct]
"""
import csv, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, H, lines = None, None, []
done_first = False
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        H = r; continue
    if H is None or len(r) < len(H):
        continue
    if r[2] != '-':
        continue            # SASS row
    lines.append((fname, int(r[0]), r[1], int(r[H.index('# Samples')] or 0), int(r[H.index('Instructions Executed')] or 0),
                  int(r[H.index('stall_long_sb')] or 0)))
ts = sum(l[3] for l in lines) or 1
ti = sum(l[4] for l in lines) or 1
print('%s: %d samples, %d warp instructions' % (kern, ts, ti))
for f, n, src, s, i, lsb in lines:
    if s * 100.0 / ts >= minpct or i * 100.0 / ti >= minpct:
        print('%-18s %5d  smp %5.1f%%  ins %5.1f%%  lsb %4.1f%%  %s' % (f, n, s * 100.0 / ts, i * 100.0 / ti, lsb * 100.0 / ts, src.strip()[:110]))
