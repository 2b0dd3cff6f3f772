"""Tuning sweep of the tcgen05 attention kernel variants (softmax warps / S prefetch / warp-level arrive) at the
64x64-level shape of the bench workload.  Each variant runs in its own process (the variant is latched at first launch)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {0: 'nw4', 1: 'nw8', 2: 'nw8+pf', 3: 'nw16', 4: 'nw16+pf', 5: 'nw8+wa', 6: 'nw8+pf+wa', 7: 'nw16+pf+wa', 8: 'nw4+wa', 9: 'pingpong', 10: 'pp+poly1/4', 11: 'pp+poly1/2', 12: 'pp16', 13: 'pp16+poly1/4',
         20: 'pt16', 21: 'pt16+poly1/8', 23: 'pt16+poly1/4', 26: 'pt8', 27: 'pt8+poly1/4',
         40: 'pt16 nosub', 41: 'pt16 nosub p1/8', 35: 'pt16 nosub p1/6', 38: 'pt16 nosub p1/12', 42: 'pt16 nosub p1/4', 43: 'pt16 nosub p1/3', 46: 'pt8 nosub', 47: 'pt8 nosub p1/4', 39: 'pt8 nosub p1/2', 48: 'pt8 nosub p1/3', 49: 'pt8 nosub p1/5', 44: 'pt8 nosub p1/6', 45: 'pt8 nosub p1/8',
         53: 'pt16 nosub wa', 54: 'pt16 nosub wa p1/8', 55: 'pt16 nosub wa p1/4', 56: 'pt8 nosub wa p1/4', 57: 'ABL wa neither', 58: 'pt16 wa pre p1/8', 59: 'pt16 pre p1/8', 64: 'pt16 pre p1/4', 65: 'pt16 pre p1/6', 66: 'pt16 pre', 67: 'pt16 pre p1/16', 68: 'pt16 pre p1/3', 69: 'ABL pre no-MUFU', 70: 'pt16 pre p1/5', 71: 'pt16 pre late p1/8', 72: 'pt16 pre late p1/4', 73: 'pt16 pre late p1/6', 74: 'pt16 pre late', 75: 'pt16 late p1/8', 76: 'ABL pre no-LDTM', 77: 'ABL pre neither', 78: 'pt16 pre x64 p1/6', 79: 'pt16 pre seq p1/6', 80: 'pt16 pre x64 late p1/6', 82: 'pt16 pre order p1/6', 83: 'pt16 pre order p1/8', 84: 'pt16 pre order', 85: 'pt16 pre order p1/4', 86: 'pt16 pre order late p1/6', 88: 'pt16 pre order p1/3', 90: 'persistent p1/6', 91: 'persistent p1/8', 92: 'persistent p1/4', 93: 'persistent', 94: 'ABL persistent no-MUFU', 60: 'pt16 wa pch p1/8', 61: 'pt16 wa pch pre p1/8', 62: 'pt16 pch p1/8', 63: 'pt8 wa pch p1/4',
         50: 'ABL no-MUFU', 51: 'ABL no-LDTM', 52: 'ABL neither',
         30: 'pt16+order', 31: 'pt16+order+p1/8', 32: 'pt16+order+p1/5', 33: 'pt16+order+p1/4', 34: 'pt16+order+p1/3', 36: 'pt8+order', 37: 'pt8+order+p1/4'}
CODE = '''
import sys, torch
sys.path.insert(0, %r)
import torch.nn.functional as F
from bench import time_attention_kernel
from imagen_pytorch_b200 import _lib
dev = torch.device('cuda')
ms = time_attention_kernel(32, dev, iters=5)
# correctness vs the mma.sync kernel on small problems: last tile with 39 keys (dead second half), 100 keys (partial second half), full
md = 0.0
for nk in (256 + 39 + 128, 256 + 100, 384, 512 + 64, 257):
    B, rows = 2, 8 * 256 + 64
    q = (F.normalize(torch.randn(B, rows, 64, device=dev), dim=-1) * 8 * 1.4426950408889634).to(torch.bfloat16)
    k = F.normalize(torch.randn(B, nk, 64, device=dev), dim=-1).to(torch.bfloat16)
    v = torch.randn(B, nk, 64, device=dev).to(torch.bfloat16)
    o1, o2 = torch.zeros_like(q), torch.zeros_like(q)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call('b200_attention', q.data_ptr(), o1.data_ptr(), rows * 64, 0, 64, rows, k.data_ptr(), v.data_ptr(), nk * 64, 0, 64, nk, B, 1, 11.8, st)
    _lib.call('b200_attention', q.data_ptr(), o2.data_ptr(), rows * 64, 0, 64, rows, k.data_ptr(), v.data_ptr(), nk * 64, 0, 64, nk, B, 1, 0.0, st)
    torch.cuda.synchronize()
    md = max(md, (o1.float() - o2.float()).abs().max().item())
print('RESULT %%.4f ms  %%.1f TFLOP/s  maxdiff %%.4f' %% (ms, 1109.98 / ms, md))
''' % ROOT

# SWEEP_VARIANTS="20,47,47:500": variant[:barrier-wait suspend hint in ns]
only = [v.strip() for v in os.environ.get('SWEEP_VARIANTS', '').split(',') if v.strip()]
for spec in (only or [str(v) for v in sorted(NAMES)]):
    var, _, ns = spec.partition(':')
    var = int(var)
    env = dict(os.environ, B200_IMAGEN_FA_VARIANT=str(var), B200_IMAGEN_FA_WAIT_NS=ns or '0')
    try:
        out = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True, timeout=300)
        res = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
        print(f'variant {spec:8s} {NAMES.get(var, "?"):16s}', res[0] if res else ('FAILED ' + out.stderr[-300:]))
    except subprocess.TimeoutExpired:
        print(f'variant {var} {NAMES[var]:12s} TIMEOUT')
