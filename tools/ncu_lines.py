"""Aggregate an ncu source page (csv) by source line: stall samples, instructions, excess shared wavefronts.  Dev tool.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
cur, hdr = None, None
for r in csv.reader(raw.splitlines()):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        ln = int(d['Line No'])
    except ValueError:
        continue

    def f(key):
        try:
            return float(d[key].replace(',', ''))
        except (KeyError, ValueError):
            return 0.0
    a = agg[(cur, ln)]
    a[0] += f('Warp Stall Sampling (All Samples)')
    a[1] += f('Instructions Executed')
    a[2] += f('L1 Wavefronts Shared Excessive')
ts, ti = sum(v[0] for v in agg.values()) or 1.0, sum(v[1] for v in agg.values()) or 1.0
print(f'total stall samples {ts:.0f}, warp instructions {ti:.0f}')
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{key[0]}:{key[1]:<5d} samples {100 * v[0] / ts:5.1f}%  instructions {100 * v[1] / ti:5.1f}%  excess shared wavefronts {v[2]:.0f}')
