"""Warm per-call timing of one U-Net evaluation (CUDA events around every C-ABI call, caches as the previous call left them).
usage: python tools/op_times.py [bs]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
im = build_model(128, 4, torch.device('cuda'))
torch.manual_seed(0)
im.sample(text_embeds=torch.randn(bs, 256, 768, device='cuda'), cond_scale=3., use_tqdm=False)
plan = next(iter(im.unets[0]._plans.values()))
plan.slots.zero_()   # sample() left the device-side schedule slot past the last table row
for _ in range(2):
    plan.launch()
runs = [plan.launch_timed() for _ in range(5)]
names = [n for n, _ in runs[0]]
med = [sorted(r[i][1] for r in runs)[2] for i in range(len(names))]
gd = iter(plan.describe_gemms())
agg = collections.defaultdict(lambda: [0, 0.0])
print(f'total {sum(med):.1f} us over {len(names)} calls')
for n, t in zip(names, med):
    key = n
    if n == 'b200_conv_gemm':
        d = next(gd)
        M = d['B'] * d['H'] * d['W']
        key = f"gemm {'conv' if d['nseg'] > 1 else 'lin '} M={M} N={d['N']} K={d['K']}" + (f" norm{d.get('norm1', 0)}{d.get('norm2', 0)}" if d.get('norm1') or d.get('norm2') else '')
    agg[key][0] += 1
    agg[key][1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f'{t:9.1f} us  x{c:3d}  {k}')
by = collections.defaultdict(float)
for n, t in zip(names, med):
    by[n] += t
print('--- by entry point')
for k, t in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f'{t:9.1f} us  {k}')
