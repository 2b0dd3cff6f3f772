"""Tuning sweep of the kept tcgen05 attention kernel variants (see the dispatch in csrc/attention_tc.cu) at the
64x64-level shape of the bench workload.  Each variant runs in its own process (the variant is latched at first launch)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {-1: 'product path', 12: 'pp16 (r01 default)', 20: 'pt16', 41: 'pt16 nosub p1/8', 60: 'pt16 wa pch p1/8', 59: 'pt16 pre p1/8',
         65: 'pt16 pre p1/6', 82: 'pt16 pre order p1/6', 112: 'split-S p1/4', 118: 'split-S late p1/4', 69: 'ABL pre no-MUFU', 114: 'ABL split-S no-MUFU',
         100: 'dual p1/6', 104: 'ABL dual no-MUFU', 90: 'persistent p1/6', 121: 'persistent split-S p1/4', 123: 'ABL persistent split-S no-MUFU'}
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
    env = dict(os.environ, B200_IMAGEN_FA_WAIT_NS=ns or '100')
    if var >= 0:
        env['B200_IMAGEN_FA_VARIANT'] = str(var)
    try:
        out = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True, timeout=300)
        res = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
        print(f'variant {spec:8s} {NAMES.get(var, "?"):16s}', res[0] if res else ('FAILED ' + out.stderr[-300:]))
    except subprocess.TimeoutExpired:
        print(f'variant {var} {NAMES[var]:12s} TIMEOUT')
