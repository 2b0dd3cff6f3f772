"""Barrier time line of one CTA of the tcgen05 self-attention kernel (development tool).  Runs the 64x64-level launch with the tracing
variant (B200_IMAGEN_FA_VARIANT=81: two issuer threads, 115: three) and prints, per key tile, when each warp role passed its wait points (cycles relative to the
tile's first event).  Roles: warp 0 TMA producer, warps 1-2 MMA issuers of query tile A/B, warps 4-19 softmax (group = (w-4)//8)."""
import os
import struct
import sys

path = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/attn_trace.bin')
os.environ.setdefault('B200_IMAGEN_FA_VARIANT', '81')
os.environ['B200_IMAGEN_FA_TRACE'] = path

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagen_pytorch_b200 import _lib  # noqa: E402

dev = torch.device('cuda')
R, rows, nk = 32, 8 * 4096, 4096 + 39
q = (F.normalize(torch.randn(R, rows, 64, device=dev), dim=-1) * 8 * 1.4426950408889634).to(torch.bfloat16)
k = F.normalize(torch.randn(R, nk, 64, device=dev), dim=-1).to(torch.bfloat16)
v = torch.randn(R, nk, 64, device=dev).to(torch.bfloat16)
o = torch.empty_like(q)
st = torch.cuda.current_stream(dev)
for _ in range(3):
    _lib.call('b200_attention', q.data_ptr(), o.data_ptr(), rows * 64, 0, 64, rows, k.data_ptr(), v.data_ptr(), nk * 64, 0, 64, nk, R, 1, 8 * 1.4426950408889634 * 1.02, st.cuda_stream)
torch.cuda.synchronize()
raw = open(path, 'rb').read()
t = struct.unpack(f'{len(raw) // 8}q', raw)
get = lambda w, j, s: t[((w * 64 + j) << 3) + s]
t0 = min(x for x in t if x > 0)
print(f'CTA {os.environ.get("B200_IMAGEN_FA_TRACE_CTA", "1000")}: first stamp = 0, last = {max(t) - t0} cycles')
print('softmax slots: 0 loop top, 1 S ready, 2 scores in registers, 3/4 before/after P-buffer wait, 5 exps issued, 6 P handed over')
print('issuer slots: 0 before K/V wait, 1 K/V ready, 2 S buffer free (S MMA issued next), 3/4 P half 0/1 ready (PV issued next), 7 end')
for j in list(range(0, 4)) + list(range(14, 18)) + list(range(30, 33)):
    print(f'--- key tile {j}')
    for w, name in ((0, 'producer'), (1, 'issuer A'), (2, 'issuer B'), (4, 'softmax A sub0 q0'), (8, 'softmax A sub1 q0'), (12, 'softmax B sub0 q0'), (16, 'softmax B sub1 q0'), (7, 'softmax A sub0 q3')):
        stamps = [get(w, j, s) for s in range(8)]
        print(f'  {name:18s}', ' '.join(f'{(x - t0):7d}' if x > 0 else '      -' for x in stamps))
# per-tile period of each softmax warp and where it goes
for w in (4, 8, 12, 16):
    per = [get(w, j + 1, 0) - get(w, j, 0) for j in range(4, 30)]
    seg = [[get(w, j, s + 1) - get(w, j, s) for j in range(4, 30)] for s in range(6)]
    print(f'warp {w}: period {sum(per) / len(per):.0f} cycles = ' + ' + '.join(f'{sum(x) / len(x):.0f}' for x in seg) + '  (S wait, TMEM load, ->P wait, P wait, exps, store+arrive)')
for w in (1, 2):
    per = [get(w, j + 1, 0) - get(w, j, 0) for j in range(4, 30)]
    seg = [[get(w, j, b) - get(w, j, a) for j in range(4, 30)] for a, b in ((0, 1), (1, 2), (2, 3), (3, 4))]
    print(f'issuer {w}: period {sum(per) / len(per):.0f} cycles = ' + ' + '.join(f'{sum(x) / len(x):.0f}' for x in seg) + '  (K/V wait, S-free wait, issue S + wait P0, issue PV0 + wait P1)')
