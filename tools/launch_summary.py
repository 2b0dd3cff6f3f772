"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST denoising step."""
import collections
import csv
import re
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
# find step boundary: last two ddpm_step_kernel launches
idx = [i for i, r in enumerate(rows) if 'ddpm_step_kernel' in r['Kernel Name']]
step = rows[idx[-2] + 1: idx[-1] + 1]
tot = sum(float(r['Metric Value'].replace(',', '')) for r in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    n = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    agg[n][0] += 1
    agg[n][1] += float(r['Metric Value'].replace(',', ''))
print(f'launches in step: {len(step)}  total {tot / 1e3:.1f} us (serialised, cold-cache)')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t / 1e3:10.1f} us {100 * t / tot:5.1f}%  x{n:3d}  {k[:100]}')
if len(sys.argv) > 3:
    print('--- launches > %s us' % sys.argv[3])
    for i, r in enumerate(step):
        t = float(r['Metric Value'].replace(',', '')) / 1e3
        if t > float(sys.argv[3]):
            print(i, f'{t:8.1f}', r['Grid Size'], re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')[:70])
