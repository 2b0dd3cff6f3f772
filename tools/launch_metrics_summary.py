"""Per-kernel totals of ONE denoising step from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...,sm__inst_executed_pipe_xu...,smsp__issue_active... --csv` launch list (tools/gpu_r2_c.sh):
time, DRAM bytes (cold-cache: ncu flushes L2 between replays), achieved DRAM GB/s, time-weighted tensor / XU pipe utilisation.
usage: python tools/launch_metrics_summary.py gpurun_out/c_launches.csv"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith('==')]
by = collections.OrderedDict()
for x in csv.DictReader(rows):
    by.setdefault(x['ID'], {'name': x['Kernel Name'], 'grid': x['Grid Size']})[x['Metric Name']] = float(x['Metric Value'].replace(',', ''))
L = list(by.values())
idx = [i for i, l in enumerate(L) if 'ddpm_step_kernel' in l['name']]
step = L[idx[-2] + 1: idx[-1] + 1]
T, TP, XU = 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'
tot = sum(l[T] for l in step) / 1e3
print(f'launches in step: {len(step)}  total {tot:.1f} us (serialised, cold-cache)')
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for l in step:
    n = re.sub(r'\(.*', '', l['name']).replace('void ', '').replace('<unnamed>::', '')
    a = agg[n]
    a[0] += 1
    a[1] += l[T] / 1e3
    a[2] += (l['dram__bytes_read.sum'] + l['dram__bytes_write.sum']) / 1e6
    a[3] += l.get(TP, 0.) * l[T] / 1e3
    a[4] += l.get(XU, 0.) * l[T] / 1e3
print(f'{"us":>9} {"share":>6} {"n":>4} {"DRAM MB":>9} {"GB/s":>7} {"tensor%":>8} {"xu%":>6}  kernel')
for k, (c, t, mb, tp, xu) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:9.1f} {100 * t / tot:5.1f}% {c:4d} {mb:9.1f} {mb / t * 1e3:7.0f} {tp / t:8.1f} {xu / t:6.1f}  {k[:100]}')
