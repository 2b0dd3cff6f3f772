"""Per-CTA fixed cost vs per-key-tile cost of the tcgen05 self-attention kernel: time the 64x64-level launch shape (32 x 32768 query
rows, 4096 CTAs of 256 rows) at several key counts and fit  t_cta = fixed + tiles * per_tile  (1 CTA per SM, 148 SMs)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagen_pytorch_b200 import _lib  # noqa: E402

dev = torch.device('cuda')
R, rows = 32, 8 * 4096
q = (F.normalize(torch.randn(R, rows, 64, device=dev), dim=-1) * 8 * 1.4426950408889634).to(torch.bfloat16)
o = torch.empty_like(q)
st = torch.cuda.current_stream(dev)
pts = []
for tiles in (1, 2, 4, 8, 16, 32, 64):
    nk = 128 * tiles + 39
    k = F.normalize(torch.randn(R, nk, 64, device=dev), dim=-1).to(torch.bfloat16)
    v = torch.randn(R, nk, 64, device=dev).to(torch.bfloat16)
    args = (q.data_ptr(), o.data_ptr(), rows * 64, 0, 64, rows, k.data_ptr(), v.data_ptr(), nk * 64, 0, 64, nk, R, 1, 8 * 1.4426950408889634 * 1.02, st.cuda_stream)
    for _ in range(3):
        _lib.call('b200_attention', *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        _lib.call('b200_attention', *args)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    waves = (R * rows / 256) / 148
    us_cta = ms * 1e3 / waves
    pts.append((tiles + 1, us_cta))
    print(f'keys {nk:5d} ({tiles + 1:2d} key tiles)  {ms:8.4f} ms  {us_cta:7.2f} us per CTA  {us_cta * 1965 / (tiles + 1):7.0f} cycles per key tile (both query tiles)')
n = len(pts)
sx, sy = sum(p[0] for p in pts), sum(p[1] for p in pts)
sxx, sxy = sum(p[0] ** 2 for p in pts), sum(p[0] * p[1] for p in pts)
slope = (n * sxy - sx * sy) / (n * sxx - sx * sx)
icpt = (sy - slope * sx) / n
print(f'fit: {icpt:.2f} us fixed per CTA + {slope:.3f} us ({slope * 1965:.0f} cycles at 1965 MHz) per 128-key tile; tensor floor 1024 cycles, MUFU floor 2048 x (1 - poly share)')
