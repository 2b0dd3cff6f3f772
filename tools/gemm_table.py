"""Join the ncu launch list of one step with the plan's GEMM descriptors (tools/profile_step.py writes both).
usage: python tools/gemm_table.py gpurun_out/launches.csv gpurun_out/plan_gemms.json"""
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
idx = [i for i, r in enumerate(rows) if 'ddpm_step_kernel' in r['Kernel Name']]
step = rows[idx[-2] + 1: idx[-1] + 1]
gemms = [r for r in step if 'conv_gemm_tc' in r['Kernel Name']]
desc = json.load(open(sys.argv[2]))
assert len(gemms) == len(desc), (len(gemms), len(desc))
tot = 0.0
by_level = {}
print('  #    us   TFLOP/s  MB/us   B  H   W     N     K  nseg  kernel')
for i, (r, d) in enumerate(zip(gemms, desc)):
    us = float(r['Metric Value'].replace(',', '')) / 1e3
    M = d['B'] * d['H'] * d['W']
    fl = 2.0 * M * d['N'] * d['K']
    npad = -(-d['N'] // 128) * 128 if d['N'] > 64 else (32 if d['N'] <= 32 else 64)
    bn = npad if npad <= 64 else (256 if '<256' in r['Kernel Name'] else 128)
    pair = 'tc2_kernel' in r['Kernel Name']
    traffic = (-(-M // 128)) * (npad // bn) * (128 + (bn // 2 if pair else bn)) * d['K'] * 2     # operand bytes through L2 -> SM
    kind = f"{d['H']}x{d['W']}" if d['nseg'] > 1 else 'linear/1x1'
    by_level.setdefault(kind, [0.0, 0])
    by_level[kind][0] += us
    by_level[kind][1] += 1
    tot += us
    print(f"{i:3d} {us:6.1f} {fl / us / 1e6:8.0f} {traffic / us / 1e6:6.2f} {d['B']:3d} {d['H']:3d} {d['W']:6d} {d['N']:5d} {d['K']:5d} {d['nseg']:4d}  "
          + r['Kernel Name'].split('conv_gemm_tc')[1][:18])
print(f'total {tot:.1f} us')
for k, (t, n) in sorted(by_level.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:12s} {t:8.1f} us  x{n}')
