"""Dev tool: per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_list.py launches.csv out.txt "header line"
"""
import csv
import sys

src, out = sys.argv[1], sys.argv[2]
head = sys.argv[3] if len(sys.argv) > 3 else ''
rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, im, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
iu = hdr.index('Metric Unit')
tot = {}
n = 0
for r in rows:
    if r is hdr or len(r) <= iv or r[im] != 'gpu__time_duration.sum':
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[iu], 1e-6)
    name = r[ik].split('(')[0][:64]
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v
    t[1] += 1
    n += 1
total = sum(t[0] for t in tot.values())
with open(out, 'w') as f:
    if head:
        f.write('# ' + head + '\n')
    f.write(f'# total {total:.1f} ms over {n} launches (serialised, cold-cache per-launch times: compare SHARES)\n')
    for name, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
        f.write(f'{ms:10.3f} ms  {100 * ms / total:5.1f} %  {c:6d}  {name}\n')
print('wrote', out)
