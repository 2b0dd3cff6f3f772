"""Thin, allocation-explicit Python wrappers over the C-ABI (used by UnetPlan and by the kernel tests)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Src, Seg, Epilogue, RowChain

BF16 = torch.bfloat16


def ceil_to(a, b):
    return (a + b - 1) // b * b


def pack_weight(mats, N, device):
    """mats: list of [N, C_seg] fp32 (one per K segment) -> bf16 [Npad, sum ceil64(C_seg)], K-major, zero padded."""
    Np = _lib.npad(N)
    cols = [ceil_to(m.shape[1], 64) for m in mats]
    out = torch.zeros(Np, sum(cols), dtype=torch.float32, device=device)
    o = 0
    for m, c in zip(mats, cols):
        out[:N, o:o + m.shape[1]] = m
        o += c
    return out.to(BF16).contiguous()


def conv_taps(k):
    r = k // 2
    return [(dh, dw) for dh in range(-r, r + 1) for dw in range(-r, r + 1)]


def conv_segments(Wt, split):
    """k x k conv (pad k//2) over a channel concat: returns (segs, mats) in the K order the kernel walks."""
    k = Wt.shape[-1]
    segs, mats = [], []
    for (dh, dw) in conv_taps(k):
        o = 0
        for i, c in enumerate(split):
            segs.append((i, dh, dw))
            mats.append(Wt[:, o:o + c, dh + k // 2, dw + k // 2])
            o += c
    return segs, mats


class GemmCall:
    """Prebuilt argument pack of one b200_conv_gemm launch; call(stream) enqueues it."""

    def __init__(self, srcs, segs, grid, wpacked, N, out_ptr, *, bias=None, act=0, out_scale=1.0, residual=None, ldr=0,
                 out_mode=_lib.OUT_BF16, ldc=0, out2_ptr=None, ldc2=0, split_col=0, rows_per_group=0, group_stride=0, row_offset=0,
                 l2_cols=0, l2_scale=None, ps_C=0, dup_rows=0, impl=_lib.IMPL_TCGEN05, scratch_ptr=None,
                 norm1=0, norm1_g=None, norm2=0, norm2_g=None, film_ptr=None, film_ld=0, rows_per_sample=0, out_norm_ptr=None, ld_norm=0):
        """srcs: list of (ptr, C, ld); segs: list of (src, dh, dw); grid: (B, H, W) of the output pixel rows."""
        self.lib = _lib.load()
        self.keep = [wpacked, bias, l2_scale]
        self.sa = (Src * len(srcs))(*[Src(p, c, ld) for (p, c, ld) in srcs])
        self.ga = (Seg * len(segs))(*[Seg(*g) for g in segs])
        e = Epilogue()
        e.bias = bias.data_ptr() if bias is not None else None
        e.act, e.out_scale = act, out_scale
        e.residual, e.ldr = residual, ldr
        e.out_mode, e.out, e.ldc = out_mode, out_ptr, ldc       # out_ptr may be None when only the fused norm output is wanted
        e.out2, e.ldc2, e.split_col = out2_ptr, ldc2, split_col
        e.rows_per_group, e.group_stride, e.row_offset = rows_per_group, group_stride, row_offset
        e.l2_cols = l2_cols
        e.l2_scale = l2_scale.data_ptr() if l2_scale is not None else None
        e.ps_C, e.dup_rows = ps_C, dup_rows
        self.e = e
        self.N, self.impl = N, impl
        self.desc = {'B': grid[0], 'H': grid[1], 'W': grid[2], 'N': N, 'nseg': len(segs), 'C': [c for (_, c, _) in srcs],
                     'K': sum(-(-srcs[g[0]][1] // 64) * 64 for g in segs), 'act': act, 'out_mode': out_mode}
        self.args = (self.sa, len(srcs), self.ga, len(segs), grid[0], grid[1], grid[2], wpacked.data_ptr(), N, C.byref(e), impl, scratch_ptr)
        if norm1:
            self.set_norm1(norm1_g)
        if norm2:
            self.set_norm2(norm2, norm2_g, out_norm_ptr, ld_norm, film_ptr=film_ptr, film_ld=film_ld, rows_per_sample=rows_per_sample)

    # ---- per-row norms fused into the epilogue (b200_epilogue ABI v2).  The struct is passed by reference at every launch, so a
    # consumer discovered later by the plan compiler can still attach its norm to an already recorded producer GEMM.
    def norm_capable(self):
        e = self.e
        return (self.impl == _lib.IMPL_TCGEN05 and 64 <= self.N <= 256 and self.N % 32 == 0 and e.out_mode == _lib.OUT_BF16 and e.split_col == 0 and
                e.rows_per_group == 0 and e.l2_cols == 0 and e.dup_rows == 0 and self.desc.get('ksplit', 1) == 1)

    def set_norm1(self, g):
        assert self.norm_capable() and self.e.norm1 == 0
        self.keep.append(g)
        self.e.norm1, self.e.norm1_g = _lib.NORM_LN, g.data_ptr()
        self.desc['norm1'] = 1

    def set_norm2(self, kind, g, out_ptr, ld, *, film_ptr=None, film_ld=0, rows_per_sample=0):
        assert self.norm_capable() and self.e.norm2 == 0
        self.keep.append(g)
        e = self.e
        e.norm2, e.norm2_g, e.out_norm, e.ld_norm = kind, g.data_ptr(), out_ptr, ld
        e.film, e.film_ld, e.rows_per_sample = film_ptr, film_ld, rows_per_sample
        self.desc['norm2'] = kind

    def __call__(self, stream):
        _lib.check(self.lib.b200_conv_gemm(*self.args, stream), 'b200_conv_gemm')


class RowChainCall:
    """Prebuilt b200_row_chain launch (gate -> LayerNorm -> + residual -> raw out; optional second norm -> out_norm).  Like GemmCall it can
    take a consumer's norm later (the descriptor is read at launch time)."""

    def __init__(self, x_ptr, ldx, M, Cc, *, gate=None, rows_per_sample=0, norm1_g=None, residual_ptr=None, ldr=0, out_ptr=None, ldo=0):
        self.lib = _lib.load()
        self.keep = [gate, norm1_g]
        e = RowChain()
        e.x, e.ldx, e.M, e.C = x_ptr, ldx, M, Cc
        e.gate = gate.data_ptr() if gate is not None else None
        e.rows_per_sample = rows_per_sample
        e.norm1 = _lib.NORM_LN if norm1_g is not None else 0
        e.norm1_g = norm1_g.data_ptr() if norm1_g is not None else None
        e.residual, e.ldr = residual_ptr, ldr
        e.out, e.ldo = out_ptr, ldo
        self.e = e
        self.args = (C.byref(e),)

    def norm_capable(self):
        return True

    def set_norm2(self, kind, g, out_ptr, ld, *, film_ptr=None, film_ld=0, rows_per_sample=0):
        assert self.e.norm2 == 0
        self.keep.append(g)
        e = self.e
        e.norm2, e.norm2_g, e.out_norm, e.ld_norm = kind, g.data_ptr(), out_ptr, ld
        e.film, e.film_ld = film_ptr, film_ld
        if rows_per_sample:
            assert e.rows_per_sample in (0, rows_per_sample)
            e.rows_per_sample = rows_per_sample

    def __call__(self, stream):
        _lib.check(self.lib.b200_row_chain(*self.args, stream), 'b200_row_chain')


def padded_bias(bias, N, device):
    bp = torch.zeros(_lib.npad(N), dtype=torch.float32, device=device)
    bp[:N] = bias
    return bp
