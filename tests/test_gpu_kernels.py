"""GPU (-m gpu): every kernel of libb200imagen.so through the C-ABI against a plain torch fp32 reference of
the same op on the same bf16-rounded inputs.  Tolerances are written next to each check."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from imagen_pytorch_b200 import _lib, ops
from imagen_pytorch_b200._lib import Src, TimeRowJob

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = 'cuda'


def stream():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def assert_close(got, ref, rtol, atol, what=''):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).float().mean().item()
    assert bad == 0, f'{what}: {bad * 100:.3f}% elements out of tolerance, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}'


# ------------------------------------------------------------------------------------------------ conv / linear

def run_conv(srcs, Wt, bias, impl, *, act=0, residual=None, out_mode=_lib.OUT_BF16, ps_C=0, dup=False, l2_cols=0, l2_scale=None,
             split_col=0, remap=None, splitk=False):
    """srcs: list of [B,H,W,C] bf16; Wt [N, sumC, k, k] fp32."""
    B, H, W = srcs[0].shape[:3]
    N = Wt.shape[0]
    M = B * H * W
    segs, mats = ops.conv_segments(Wt, [s.shape[-1] for s in srcs])
    wp = ops.pack_weight(mats, N, DEV)
    reps = 2 if dup else 1
    if out_mode == _lib.OUT_BF16:
        rows = M * reps if remap is None else remap[3]
        out = torch.zeros(rows, N if split_col == 0 else split_col, dtype=BF16, device=DEV)
        ldc = out.shape[1]
    elif out_mode == _lib.OUT_PIXEL_SHUFFLE:
        out = torch.zeros(B * 4 * H * W, ps_C, dtype=BF16, device=DEV)
        ldc = ps_C
    elif out_mode == _lib.OUT_F32_NCHW:
        out = torch.zeros(B, N, H, W, dtype=torch.float32, device=DEV)
        ldc = 0
    else:
        out = torch.zeros(M * reps, N, dtype=torch.float32, device=DEV)
        ldc = N
    out2 = torch.zeros(out.shape[0], N - split_col, dtype=BF16, device=DEV) if split_col else None
    scratch = torch.empty(M * _lib.npad(N), dtype=torch.float32, device=DEV) if impl == 1 else None
    if impl == 0 and splitk:
        ks = _lib.load().b200_conv_gemm_splitk(B, H, W, N, sum(-(-srcs[g[0]].shape[-1] // 64) * 64 for g in segs))
        scratch = torch.empty(ks * M * _lib.npad(N), dtype=torch.float32, device=DEV) if ks > 1 else None
    kw = {}
    if remap is not None:
        kw = dict(rows_per_group=remap[0], group_stride=remap[1], row_offset=remap[2])
    call = ops.GemmCall([(s.data_ptr(), s.shape[-1], s.shape[-1]) for s in srcs], segs, (B, H, W), wp, N, out.data_ptr(),
                        bias=ops.padded_bias(bias, N, DEV) if bias is not None else None, act=act,
                        residual=residual.data_ptr() if residual is not None else None, ldr=residual.shape[-1] if residual is not None else 0,
                        out_mode=out_mode, ldc=ldc, out2_ptr=out2.data_ptr() if out2 is not None else None,
                        ldc2=out2.shape[1] if out2 is not None else 0, split_col=split_col, l2_cols=l2_cols, l2_scale=l2_scale, ps_C=ps_C,
                        dup_rows=M if dup else 0, impl=impl, scratch_ptr=scratch.data_ptr() if scratch is not None else None, **kw)
    call(stream())
    torch.cuda.synchronize()
    return out, out2


def ref_conv(srcs, Wt, bias, act=0, residual=None):
    x = torch.cat([s.float() for s in srcs], dim=-1).permute(0, 3, 1, 2)
    y = F.conv2d(x, Wt.to(BF16).float(), bias, padding=Wt.shape[-1] // 2).permute(0, 2, 3, 1)
    if act == _lib.ACT_SILU:
        y = F.silu(y)
    elif act == _lib.ACT_GELU:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.float().view(y.shape)
    return y


CONV_CASES = {
    'c3_16x16_64to64': dict(B=2, H=16, W=16, Cs=[64], N=64, k=3),
    'c3_8x8_128to128_ragged_batch': dict(B=3, H=8, W=8, Cs=[128], N=128, k=3),
    'c3_32x32_concat_32p96_to256': dict(B=2, H=32, W=32, Cs=[32, 96], N=256, k=3),
    'c3_64x64_128to128': dict(B=2, H=64, W=64, Cs=[128], N=128, k=3),
    'c3_24x24_nonpow2_32to32': dict(B=2, H=24, W=24, Cs=[32], N=32, k=3),
    'c1_16x16_192to512': dict(B=2, H=16, W=16, Cs=[192], N=512, k=1),
    'c3_128wide_64to64': dict(B=1, H=4, W=256, Cs=[64], N=64, k=3),
    'c3_8x8_1536to1024_deepK': dict(B=4, H=8, W=8, Cs=[1024, 512], N=1024, k=3),          # 8 tiles of 128 x 256: K split over a 4-CTA cluster
    'c3_8x8_256to1024_b32_cluster_k2': dict(B=32, H=8, W=8, Cs=[256], N=1024, k=3),       # 64 tiles: K split over CTA pairs
    'c3_8x8_512to768_b26_cluster_k3': dict(B=26, H=8, W=8, Cs=[512], N=768, k=3),         # 13 x 3 tiles: 3-CTA clusters
}


@pytest.mark.parametrize('impl', [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_CHECKER], ids=['tcgen05', 'simt'])
@pytest.mark.parametrize('name', list(CONV_CASES))
def test_conv_gemm_matches_conv2d(name, impl):
    c = CONV_CASES[name]
    srcs = [rnd(c['B'], c['H'], c['W'], cs, seed=i).to(BF16) for i, cs in enumerate(c['Cs'])]
    K = sum(c['Cs']) * c['k'] ** 2
    Wt = rnd(c['N'], sum(c['Cs']), c['k'], c['k'], scale=1 / math.sqrt(K), seed=7)
    bias = rnd(c['N'], scale=0.1, seed=9)
    out, _ = run_conv(srcs, Wt, bias, impl)
    ref = ref_conv(srcs, Wt, bias).reshape(-1, c['N'])
    # bf16 output rounding (2^-9 relative) + fp32 accumulation-order noise
    assert_close(out, ref, rtol=8e-3, atol=8e-3, what=name)


@pytest.mark.parametrize('shape', [(8, 8, 8, [512], 512, 3), (4, 8, 8, [256, 128], 256, 3), (2, 1, 100, [2048], 512, 1)])
def test_conv_gemm_with_split_k_workspace(shape):
    """Few row tiles x long K (the 8x8 levels): with a workspace the library may split K over two CTAs per 128 x 256 tile and finish
    with a summing epilogue kernel (B200_IMAGEN_GEMM_SPLITK=1; otherwise this runs the unsplit path).  Same result either way."""
    B, H, W, Cs, N, k = shape
    srcs = [rnd(B, H, W, cs, seed=i).to(BF16) for i, cs in enumerate(Cs)]
    K = sum(Cs) * k * k
    Wt = rnd(N, sum(Cs), k, k, scale=1 / math.sqrt(K), seed=7)
    bias = rnd(N, scale=0.1, seed=9)
    res = rnd(B * H * W, N, seed=5).to(BF16)
    out, _ = run_conv(srcs, Wt, bias, _lib.IMPL_TCGEN05, act=_lib.ACT_SILU, residual=res, splitk=True)
    ref = ref_conv(srcs, Wt, bias, act=_lib.ACT_SILU, residual=res.view(B, H, W, N)).reshape(-1, N)
    assert_close(out, ref, rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize('impl', [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_CHECKER], ids=['tcgen05', 'simt'])
def test_linear_ragged_rows_gelu_residual(impl):
    M, K, N = 300, 192, 320
    x = rnd(1, 1, M, K).to(BF16)
    Wt = rnd(N, K, 1, 1, scale=1 / math.sqrt(K), seed=3)
    res = rnd(M, N, seed=5).to(BF16)
    out, _ = run_conv([x], Wt, None, impl, act=_lib.ACT_GELU, residual=res)
    ref = ref_conv([x], Wt, None, act=_lib.ACT_GELU, residual=res.view(1, 1, M, N)).reshape(M, N)
    assert_close(out, ref, rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize('impl', [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_CHECKER], ids=['tcgen05', 'simt'])
def test_final_conv_fp32_nchw_and_tiny_m_fp32_rows(impl):
    srcs = [rnd(2, 16, 16, 32).to(BF16), rnd(2, 16, 16, 8, seed=4).to(BF16)]
    Wt = rnd(3, 40, 3, 3, scale=0.05, seed=2)
    bias = rnd(3, scale=0.1)
    out, _ = run_conv(srcs, Wt, bias, impl, out_mode=_lib.OUT_F32_NCHW)
    ref = ref_conv(srcs, Wt, bias).permute(0, 3, 1, 2)
    assert_close(out, ref, rtol=1e-3, atol=2e-3, what='nchw')        # fp32 output: only accumulation-order noise
    x = rnd(1, 1, 4, 128).to(BF16)
    Wl = rnd(200, 128, 1, 1, scale=0.1)
    out, _ = run_conv([x], Wl, rnd(200, scale=0.1, seed=1), impl, out_mode=_lib.OUT_F32, dup=True)
    ref = ref_conv([x], Wl, rnd(200, scale=0.1, seed=1)).reshape(4, 200)
    assert_close(out[:4], ref, rtol=1e-3, atol=2e-3, what='f32 rows')
    assert torch.equal(out[:4], out[4:])                               # dup_rows (classifier-free-guidance batch duplication)


@pytest.mark.parametrize('impl', [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_CHECKER], ids=['tcgen05', 'simt'])
def test_pixel_shuffle_epilogue(impl):
    B, H, W, Cin, Co = 2, 8, 8, 64, 32
    x = rnd(B, H, W, Cin).to(BF16)
    W_ref = rnd(4 * Co, Cin, 1, 1, scale=1 / 8, seed=1)                # reference order: out channel c'*4 + r
    b_ref = rnd(4 * Co, scale=0.1, seed=2)
    Wp = W_ref.view(Co, 4, Cin, 1, 1).permute(1, 0, 2, 3, 4).reshape(4 * Co, Cin, 1, 1)
    bp = b_ref.view(Co, 4).permute(1, 0).reshape(-1)
    out, _ = run_conv([x], Wp, bp, impl, act=_lib.ACT_SILU, out_mode=_lib.OUT_PIXEL_SHUFFLE, ps_C=Co)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), W_ref.to(BF16).float(), b_ref)
    ref = F.pixel_shuffle(F.silu(y), 2).permute(0, 2, 3, 1).reshape(-1, Co)   # PixelShuffleUpsample (imagen_pytorch.py:613-617)
    assert_close(out, ref, rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize('impl', [_lib.IMPL_TCGEN05, _lib.IMPL_SIMT_CHECKER], ids=['tcgen05', 'simt'])
def test_qk_l2norm_and_kv_scatter_epilogues(impl):
    M, Cc = 256, 96
    x = rnd(1, 1, M, Cc).to(BF16)
    Wq = rnd(512, Cc, 1, 1, scale=0.1, seed=1)
    qs = (rnd(64, seed=2).abs() + 0.5)
    out, _ = run_conv([x], Wq, None, impl, l2_cols=512, l2_scale=qs)
    q = F.linear(x.float().view(M, Cc), Wq.to(BF16).float().view(512, Cc)).view(M, 8, 64)
    ref = (F.normalize(q, dim=-1) * qs).view(M, 512)
    assert_close(out, ref, rtol=8e-3, atol=2e-3, what='q l2norm')
    # to_kv: k normalised, v raw, scattered behind 5 prefix rows of each sample (n = 64 rows per sample)
    Wkv = rnd(128, Cc, 1, 1, scale=0.1, seed=3)
    n, npre = 64, 5
    Mtot = npre + n
    out, out2 = run_conv([x], Wkv, None, impl, l2_cols=64, l2_scale=qs, split_col=64, remap=(n, Mtot, npre, (M // n) * Mtot))
    kv = F.linear(x.float().view(M, Cc), Wkv.to(BF16).float().view(128, Cc))
    k_ref = F.normalize(kv[:, :64], dim=-1) * qs
    Kb, Vb = out.view(M // n, Mtot, 64), out2.view(M // n, Mtot, 64)
    assert_close(Kb[:, npre:].reshape(M, 64), k_ref, rtol=8e-3, atol=2e-3, what='k scatter')
    assert_close(Vb[:, npre:].reshape(M, 64), kv[:, 64:], rtol=8e-3, atol=4e-3, what='v scatter')
    assert Kb[:, :npre].abs().max() == 0                              # prefix rows untouched


def _ln_ref(x, g):
    mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


# ------------------------------------------------------------------------------------------------ transposed kernel (<= 128 output channels, many rows)

@pytest.mark.parametrize('case', ['conv3x3_plain', 'conv3x3_res', 'conv3x3_rms_film_only', 'conv3x3_raw_and_ln', 'two_sources_res_rms', 'ragged_96ch_silu', 'linear_k512'])
def test_transposed_gemm_for_128_output_channels(case):
    """Shapes of conv_gemm_tcT_kernel (D^T = W X^T with the weights as the M = 128 operand and 256 pixels as the N operand, epilogue transposed
    through shared memory; Npad == 128, K >= 512, >= one 256-pixel tile per SM).  The kernel is opt-in (B200_IMAGEN_GEMM_T=1) since the
    elect.sync issue fix made the row-major kernel faster: this test runs on whichever kernel the process selects, and
    test_transposed_kernel_switch re-runs it in a child process with the switch on.  Same torch fp32 references as the row-major tests."""
    cfg = {
        'conv3x3_plain': dict(B=3, H=128, W=128, Cs=[128], N=128, k=3, res=False, n2=0, film=False, raw=True, act=0),
        'conv3x3_res': dict(B=3, H=128, W=128, Cs=[128], N=128, k=3, res=True, n2=0, film=False, raw=True, act=0),
        'conv3x3_rms_film_only': dict(B=10, H=64, W=64, Cs=[128], N=128, k=3, res=False, n2=2, film=True, raw=False, act=0),
        'conv3x3_raw_and_ln': dict(B=10, H=64, W=64, Cs=[128], N=128, k=3, res=False, n2=1, film=False, raw=True, act=0),
        'two_sources_res_rms': dict(B=10, H=64, W=64, Cs=[128, 64], N=128, k=3, res=True, n2=2, film=False, raw=True, act=0),
        'ragged_96ch_silu': dict(B=5, H=100, W=130, Cs=[64], N=96, k=3, res=False, n2=2, film=True, raw=True, act=_lib.ACT_SILU),
        'linear_k512': dict(B=1, H=1, W=40001, Cs=[512], N=128, k=1, res=True, n2=0, film=False, raw=True, act=0),
    }[case]
    B, H, W, Cs, N, k = (cfg[x] for x in ('B', 'H', 'W', 'Cs', 'N', 'k'))
    M = B * H * W
    assert M >= 256 * 148, 'too few rows to take the transposed kernel'
    srcs = [rnd(B, H, W, c, seed=10 + i).to(BF16) for i, c in enumerate(Cs)]
    Wt = rnd(N, sum(Cs), k, k, scale=1.0 / math.sqrt(sum(Cs) * k * k), seed=2)
    bias = rnd(N, scale=0.1, seed=3)
    res = (rnd(M, N, seed=4) * 1.5 + 0.3).to(BF16) if cfg['res'] else None
    g2 = (1 + 0.2 * rnd(N, seed=6)).contiguous()
    film = rnd(B, 2 * N + 8, scale=0.3, seed=7).contiguous() if cfg['film'] else None
    out = torch.zeros(M, N, dtype=BF16, device=DEV) if cfg['raw'] else None
    out_n = torch.zeros(M, N, dtype=BF16, device=DEV) if cfg['n2'] else None
    segs, mats = ops.conv_segments(Wt, Cs)
    call = ops.GemmCall([(x.data_ptr(), x.shape[-1], x.shape[-1]) for x in srcs], segs, (B, H, W), ops.pack_weight(mats, N, DEV), N,
                        out.data_ptr() if out is not None else None, bias=ops.padded_bias(bias, N, DEV), act=cfg['act'],
                        residual=res.data_ptr() if res is not None else None, ldr=N, ldc=N)
    if cfg['n2']:
        g2k = g2 * math.sqrt(N) if cfg['n2'] == 2 else g2
        call.set_norm2(cfg['n2'], g2k.contiguous(), out_n.data_ptr(), N, film_ptr=film.data_ptr() if film is not None else None,
                       film_ld=film.shape[1] if film is not None else 0, rows_per_sample=H * W)
    call(stream())
    torch.cuda.synchronize()
    x = torch.cat([s_.float() for s_ in srcs], dim=-1)
    v = F.conv2d(x.permute(0, 3, 1, 2), Wt.to(BF16).float(), bias, padding=k // 2).permute(0, 2, 3, 1).reshape(M, N)
    if cfg['act'] == _lib.ACT_SILU:
        v = F.silu(v)
    w = v + (res.float() if res is not None else 0.)
    if out is not None:
        assert_close(out, w, 1.2e-2, 1.5e-2, f'{case}: raw')
    if cfg['n2']:
        if cfg['n2'] == 1:
            y = _ln_ref(w, g2)
        else:
            y = F.normalize(w, dim=-1) * g2 * math.sqrt(N)
            if film is not None:
                fr = film.repeat_interleave(H * W, dim=0)
                y = y * (fr[:, :N] + 1) + fr[:, N:2 * N]
            y = F.silu(y)
        assert_close(out_n, y, 1.2e-2, 1.5e-2, f'{case}: normalised')


def test_transposed_kernel_switch():
    """The opt-in transposed GEMM kernel still passes its parity cases (child process: the switch is read once per process)."""
    import os
    import subprocess
    import sys
    if os.environ.get('B200_IMAGEN_GEMM_T') == '1':
        pytest.skip('already inside the child')
    env = dict(os.environ, B200_IMAGEN_GEMM_T='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-p', 'no:cacheprovider', __file__, '-m', 'gpu', '-k', 'test_transposed_gemm_for_128_output_channels'],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '7 passed' in r.stdout, r.stdout[-500:]


# ------------------------------------------------------------------------------------------------ norms fused into the GEMM epilogue

@pytest.mark.parametrize('case', ['conv3x3_rms_film_only', 'linear_ln_res_ln', 'linear_ln_res_rms_film_only', 'linear_gelu_ln_only', 'conv_res_ln_lazy',
                                  'n64_rms', 'n192_ln_res_rms', 'n256_ln_res_ln', 'ragged_rows_rms'])
def test_gemm_epilogue_fused_norms(case):
    """b200_epilogue.norm1 / norm2 (ABI v2): LayerNorm -> + residual -> {LayerNorm | RMSNorm+FiLM+SiLU} inside the tcgen05 GEMM epilogue
    against the same chain in torch fp32 on the bf16-rounded operands (imagen_pytorch.py:331-349, :322-329, :683-691)."""
    cfg = {
        'conv3x3_rms_film_only': dict(B=2, H=16, W=16, C=128, N=128, k=3, n1=False, res=False, n2=2, film=True, raw=False, act=0),
        'linear_ln_res_ln': dict(B=1, H=1, W=777, C=512, N=128, k=1, n1=True, res=True, n2=1, film=False, raw=True, act=0),
        'linear_ln_res_rms_film_only': dict(B=2, H=8, W=8, C=512, N=128, k=1, n1=True, res=True, n2=2, film=True, raw=False, act=0),
        'linear_gelu_ln_only': dict(B=1, H=1, W=1000, C=128, N=256, k=1, n1=False, res=False, n2=1, film=False, raw=False, act=_lib.ACT_GELU),
        'conv_res_ln_lazy': dict(B=3, H=8, W=8, C=64, N=128, k=3, n1=False, res=True, n2=1, film=False, raw=True, act=0),
        'n64_rms': dict(B=2, H=8, W=8, C=64, N=64, k=3, n1=False, res=False, n2=2, film=True, raw=True, act=0),
        'n192_ln_res_rms': dict(B=2, H=8, W=8, C=192, N=192, k=1, n1=True, res=True, n2=2, film=False, raw=True, act=0),
        'n256_ln_res_ln': dict(B=1, H=16, W=16, C=256, N=256, k=3, n1=True, res=True, n2=1, film=False, raw=True, act=0),
        'ragged_rows_rms': dict(B=3, H=5, W=7, C=128, N=128, k=3, n1=False, res=True, n2=2, film=True, raw=True, act=0),
    }[case]
    B, H, W, Cin, N, k = (cfg[x] for x in ('B', 'H', 'W', 'C', 'N', 'k'))
    M = B * H * W
    x = rnd(B, H, W, Cin, seed=1).to(BF16)
    Wt = rnd(N, Cin, k, k, scale=1.0 / math.sqrt(Cin * k * k), seed=2)
    bias = rnd(N, scale=0.1, seed=3)
    res = (rnd(M, N, seed=4) * 1.5 + 0.3).to(BF16) if cfg['res'] else None
    g1 = (1 + 0.2 * rnd(N, seed=5)).contiguous()
    g2 = (1 + 0.2 * rnd(N, seed=6)).contiguous()
    film = rnd(B, 2 * N + 8, scale=0.3, seed=7).contiguous() if cfg['film'] else None    # [scale | shift] rows with a row pitch > 2N
    out = torch.zeros(M, N, dtype=BF16, device=DEV) if cfg['raw'] else None
    out_n = torch.zeros(M, N, dtype=BF16, device=DEV)
    segs, mats = ops.conv_segments(Wt, [Cin])
    call = ops.GemmCall([(x.data_ptr(), Cin, Cin)], segs, (B, H, W), ops.pack_weight(mats, N, DEV), N, out.data_ptr() if out is not None else None,
                        bias=ops.padded_bias(bias, N, DEV), act=cfg['act'], residual=res.data_ptr() if res is not None else None, ldr=N, ldc=N)
    assert call.norm_capable()
    if cfg['n1']:
        call.set_norm1(g1)
    g2k = g2 * math.sqrt(N) if cfg['n2'] == 2 else g2            # the caller folds sqrt(C) into the RMS gamma
    call.set_norm2(cfg['n2'], g2k.contiguous(), out_n.data_ptr(), N, film_ptr=film.data_ptr() if film is not None else None,
                   film_ld=film.shape[1] if film is not None else 0, rows_per_sample=H * W)
    call(stream())
    torch.cuda.synchronize()
    # torch fp32 reference
    v = F.conv2d(x.float().permute(0, 3, 1, 2), Wt.to(BF16).float(), bias, padding=k // 2).permute(0, 2, 3, 1).reshape(M, N)
    if cfg['act'] == _lib.ACT_GELU:
        v = F.gelu(v)
    if cfg['n1']:
        v = _ln_ref(v, g1)
    w = v + (res.float() if res is not None else 0.)
    if cfg['n2'] == 1:
        y = _ln_ref(w, g2)
    else:
        y = F.normalize(w, dim=-1) * g2 * math.sqrt(N)
        if film is not None:
            fr = film.repeat_interleave(H * W, dim=0)
            y = y * (fr[:, :N] + 1) + fr[:, N:2 * N]
        y = F.silu(y)
    if out is not None:
        assert_close(out, w, 1.2e-2, 1.5e-2, f'{case}: raw')
    assert_close(out_n, y, 1.2e-2, 1.5e-2, f'{case}: normalised')


# ------------------------------------------------------------------------------------------------ attention

def ref_attention(q, k, v):
    """q [P, rows, 64] already scaled by 8*log2e; softmax base 2."""
    s = torch.einsum('pid,pjd->pij', q.float(), k.float()) * math.log(2)
    return torch.einsum('pij,pjd->pid', s.softmax(dim=-1), v.float())


BOUND = 8 * 1.4426950408889634 * 1.02      # unit q/k scales: |q.k| * 8 * log2e <= 11.54 (+2% bf16 slack)


@pytest.mark.parametrize('path', ['tc', 'mma'])
@pytest.mark.parametrize('rows,nkeys', [(8 * 256, 256 + 39), (8 * 64, 103), (8 * 1024, 1024 + 39), (200, 1), (8 * 4096, 4096 + 39)])
def test_multi_query_attention(rows, nkeys, path):
    B = 2
    q = (F.normalize(rnd(B, rows, 64), dim=-1) * 8 * 1.4426950408889634).to(BF16)
    k = F.normalize(rnd(B, nkeys, 64, seed=1), dim=-1).to(BF16)
    v = rnd(B, nkeys, 64, seed=2).to(BF16)
    o = torch.zeros_like(q)
    _lib.call('b200_attention', q.data_ptr(), o.data_ptr(), rows * 64, 0, 64, rows, k.data_ptr(), v.data_ptr(), nkeys * 64, 0, 64, nkeys, B, 1,
              BOUND if path == 'tc' else 0.0, stream())
    torch.cuda.synchronize()
    # P is rounded to bf16 before P@V and O to bf16 on store
    assert_close(o, ref_attention(q, k, v), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('path', ['tc', 'mma'])
@pytest.mark.parametrize('n,nk', [(100, 39), (4096, 39), (64, 259), (1000, 64), (300, 5), (20000, 39)])
def test_cross_attention_layout_per_head_kv(n, nk, path):
    B, heads = 2, 8
    inner = heads * 64
    q = (F.normalize(rnd(B, n, heads, 64), dim=-1) * 8 * 1.4426950408889634).to(BF16)
    k = F.normalize(rnd(B, nk, heads, 64, seed=1), dim=-1).to(BF16)
    v = rnd(B, nk, heads, 64, seed=2).to(BF16)
    o = torch.zeros_like(q)
    _lib.call('b200_attention', q.data_ptr(), o.data_ptr(), n * inner, 64, inner, n, k.data_ptr(), v.data_ptr(), nk * inner, 64, inner, nk, B, heads,
              BOUND if path == 'tc' else 0.0, stream())
    torch.cuda.synchronize()
    qq = q.permute(0, 2, 1, 3).reshape(B * heads, n, 64)
    kk = k.permute(0, 2, 1, 3).reshape(B * heads, nk, 64)
    vv = v.permute(0, 2, 1, 3).reshape(B * heads, nk, 64)
    ref = ref_attention(qq, kk, vv).view(B, heads, n, 64).permute(0, 2, 1, 3)
    assert_close(o, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('B,n,nk,heads', [(600, 128, 39, 8), (37, 200, 39, 4), (3, 4096, 17, 8)])
def test_cross_attention_tc_many_samples_per_cta(B, n, nk, heads):
    """Persistent tcgen05 cross-attention (cross_attn_tc_kernel): a CTA's item range spans several samples (its resident K/V tiles are
    reloaded at every sample boundary), partial last query tiles, fewer than 8 heads."""
    inner = heads * 64
    q = (F.normalize(rnd(B, n, heads, 64), dim=-1) * 8 * 1.4426950408889634).to(BF16)
    k = F.normalize(rnd(B, nk, heads, 64, seed=1), dim=-1).to(BF16)
    v = rnd(B, nk, heads, 64, seed=2).to(BF16)
    o = torch.zeros_like(q)
    _lib.call('b200_attention', q.data_ptr(), o.data_ptr(), n * inner, 64, inner, n, k.data_ptr(), v.data_ptr(), nk * inner, 64, inner, nk, B, heads,
              BOUND, stream())
    torch.cuda.synchronize()
    qq = q.permute(0, 2, 1, 3).reshape(B * heads, n, 64)
    kk = k.permute(0, 2, 1, 3).reshape(B * heads, nk, 64)
    vv = v.permute(0, 2, 1, 3).reshape(B * heads, nk, 64)
    ref = ref_attention(qq, kk, vv).view(B, heads, n, 64).permute(0, 2, 1, 3)
    assert_close(o, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('variant', [12, 60, 82, 90, 100])
def test_attention_alternative_kernels_stay_correct(variant):
    """The kept, selectable alternatives of the self-attention kernel (round-1 ping-pong kernel, chunked P hand-over, MUFU token, persistent at
    any length, dual softmax) pass the same parity cases as the product path (child process: the variant is read once per process)."""
    import os
    import subprocess
    import sys
    if os.environ.get('B200_IMAGEN_FA_VARIANT'):
        pytest.skip('already inside a child')
    env = dict(os.environ, B200_IMAGEN_FA_VARIANT=str(variant))
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-p', 'no:cacheprovider', __file__, '-m', 'gpu', '-k', 'test_multi_query_attention or test_cross_attention_layout_per_head_kv'],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout and 'no tests ran' not in r.stdout, r.stdout[-500:]


# ------------------------------------------------------------------------------------------------ row-wise kernels

@pytest.mark.parametrize('Cs,film', [([128], True), ([256, 128], False), ([1536, 768], True), ([32], False), ([40, 24], True)])
def test_rmsnorm_film_silu(Cs, film):
    B, n = 3, 50
    M, Ct = B * n, sum(Cs)
    srcs = [rnd(M, c, seed=i).to(BF16) for i, c in enumerate(Cs)]
    gamma = rnd(Ct, seed=5).abs() + 0.5
    fl = rnd(B, 2 * Ct + 16, scale=0.3, seed=6) if film else None
    out = torch.zeros(M, Ct, dtype=BF16, device=DEV)
    sa = (Src * len(srcs))(*[Src(s.data_ptr(), s.shape[1], s.shape[1]) for s in srcs])
    g = (gamma * math.sqrt(Ct)).contiguous()
    _lib.call('b200_rmsnorm_film_silu', sa, len(srcs), 2 ** -0.5, g.data_ptr(), fl.data_ptr() if film else None, fl.shape[1] if film else 0, n,
              out.data_ptr(), Ct, M, stream())
    torch.cuda.synchronize()
    x = srcs[0].float() if len(srcs) == 1 else torch.cat((srcs[0].float(), srcs[1].float() * 2 ** -0.5), dim=1)
    y = F.normalize(x, dim=1) * math.sqrt(Ct) * gamma                     # ChanRMSNorm (imagen_pytorch.py:322-329)
    if film:
        f = fl.repeat_interleave(n, dim=0)
        y = y * (f[:, :Ct] + 1) + f[:, Ct:2 * Ct]
    assert_close(out, F.silu(y), rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize('Cc,residual', [(128, True), (1024, False), (2048, True), (40, False)])
def test_layernorm_gain_only_with_residual(Cc, residual):
    M = 77
    x = rnd(M, Cc, scale=2.0).to(BF16)
    g = rnd(Cc, seed=1).abs() + 0.5
    r = rnd(M, Cc, seed=2).to(BF16) if residual else None
    out = torch.zeros(M, Cc, dtype=BF16, device=DEV)
    _lib.call('b200_layernorm', x.data_ptr(), Cc, g.data_ptr(), None, 1e-5, r.data_ptr() if residual else None, Cc, out.data_ptr(), Cc, M, Cc, stream())
    torch.cuda.synchronize()
    xf = x.float()
    ref = (xf - xf.mean(-1, keepdim=True)) * (xf.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * g   # imagen_pytorch.py:339-349
    if residual:
        ref = ref + r.float()
    assert_close(out, ref, rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize('Cs,film,pitch', [([128], True, 0), ([64, 64], True, 32), ([256], False, 0), ([256, 128], True, 0), ([16], False, 8)])
def test_rmsnorm_film_silu_many_rows(Cs, film, pitch):
    """Many rows (several waves of blocks), ragged tail, strided sources."""
    B, n = 5, 4099 * (8 if sum(Cs) <= 16 else 1)
    M, Ct = B * n, sum(Cs)
    wide = [rnd(M, c + pitch, seed=i).to(BF16) for i, c in enumerate(Cs)]
    srcs = [w[:, :c] for w, c in zip(wide, Cs)]
    gamma = rnd(Ct, seed=5).abs() + 0.5
    fl = rnd(B, 2 * Ct + 16, scale=0.3, seed=6) if film else None
    out = torch.zeros(M, Ct, dtype=BF16, device=DEV)
    sa = (Src * len(srcs))(*[Src(w.data_ptr(), c, w.shape[1]) for w, c in zip(wide, Cs)])
    g = (gamma * math.sqrt(Ct)).contiguous()
    _lib.call('b200_rmsnorm_film_silu', sa, len(srcs), 2 ** -0.5, g.data_ptr(), fl.data_ptr() if film else None, fl.shape[1] if film else 0, n,
              out.data_ptr(), Ct, M, stream())
    torch.cuda.synchronize()
    x = srcs[0].float() if len(srcs) == 1 else torch.cat((srcs[0].float(), srcs[1].float() * 2 ** -0.5), dim=1)
    y = F.normalize(x, dim=1) * math.sqrt(Ct) * gamma
    if film:
        f = fl.repeat_interleave(n, dim=0)
        y = y * (f[:, :Ct] + 1) + f[:, Ct:2 * Ct]
    assert_close(out, F.silu(y), rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize('Cc,residual,pitch', [(128, True, 0), (128, False, 64), (256, True, 8), (512, True, 0), (64, False, 0)])
def test_layernorm_many_rows(Cc, residual, pitch):
    M = 20495 * (2 if Cc >= 512 or Cc <= 64 else 1)
    xw = rnd(M, Cc + pitch, scale=2.0).to(BF16)
    x = xw[:, :Cc]
    g = rnd(Cc, seed=1).abs() + 0.5
    beta = rnd(Cc, seed=3)
    rw = rnd(M, Cc + pitch, seed=2).to(BF16) if residual else None
    out = torch.zeros(M, Cc, dtype=BF16, device=DEV)
    _lib.call('b200_layernorm', xw.data_ptr(), Cc + pitch, g.data_ptr(), beta.data_ptr(), 1e-5, rw.data_ptr() if residual else None, Cc + pitch,
              out.data_ptr(), Cc, M, Cc, stream())
    torch.cuda.synchronize()
    xf = x.float()
    ref = (xf - xf.mean(-1, keepdim=True)) * (xf.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * g + beta
    if residual:
        ref = ref + rw[:, :Cc].float()
    assert_close(out, ref, rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize('case', ['ln_res', 'gate_res', 'ln_res_ln', 'gate_res_rms_film', 'ln_res_rms_film_only', 'gate_res_ln_c40', 'ln_res_ln_c2048'])
def test_row_chain_kernel(case):
    """b200_row_chain: [gate] -> [LayerNorm] -> + residual -> raw out; [LayerNorm | RMSNorm+FiLM+SiLU] -> out_norm, in one pass over the row."""
    cfg = {
        'ln_res': dict(B=3, n=700, C=512, gate=False, n1=True, n2=0, film=False, raw=True),
        'gate_res': dict(B=3, n=1024, C=128, gate=True, n1=False, n2=0, film=False, raw=True),
        'ln_res_ln': dict(B=2, n=256, C=1024, gate=False, n1=True, n2=1, film=False, raw=True),
        'gate_res_rms_film': dict(B=4, n=300, C=256, gate=True, n1=False, n2=2, film=True, raw=True),
        'ln_res_rms_film_only': dict(B=2, n=64, C=1024, gate=False, n1=True, n2=2, film=True, raw=False),
        'gate_res_ln_c40': dict(B=3, n=77, C=40, gate=True, n1=False, n2=1, film=False, raw=True),
        'ln_res_ln_c2048': dict(B=2, n=33, C=2048, gate=False, n1=True, n2=1, film=False, raw=True),
    }[case]
    B, n, Cc = cfg['B'], cfg['n'], cfg['C']
    M = B * n
    x = rnd(M, Cc, seed=1).to(BF16)
    res = (rnd(M, Cc, seed=2) * 1.3 + 0.2).to(BF16)
    gate = torch.sigmoid(rnd(B, Cc, seed=3)).contiguous() if cfg['gate'] else None
    g1 = (1 + 0.2 * rnd(Cc, seed=4)).contiguous()
    g2 = (1 + 0.2 * rnd(Cc, seed=5)).contiguous()
    film = rnd(B, 2 * Cc + 8, scale=0.3, seed=6).contiguous() if cfg['film'] else None
    out = torch.zeros(M, Cc, dtype=BF16, device=DEV) if cfg['raw'] else None
    out_n = torch.zeros(M, Cc, dtype=BF16, device=DEV) if cfg['n2'] else None
    call = ops.RowChainCall(x.data_ptr(), Cc, M, Cc, gate=gate, rows_per_sample=n, norm1_g=g1 if cfg['n1'] else None, residual_ptr=res.data_ptr(), ldr=Cc,
                            out_ptr=out.data_ptr() if out is not None else None, ldo=Cc)
    if cfg['n2']:
        g2k = (g2 * math.sqrt(Cc)).contiguous() if cfg['n2'] == 2 else g2
        call.set_norm2(cfg['n2'], g2k, out_n.data_ptr(), Cc, film_ptr=film.data_ptr() if film is not None else None,
                       film_ld=film.shape[1] if film is not None else 0, rows_per_sample=n)
    call(stream())
    torch.cuda.synchronize()
    v = x.float()
    if gate is not None:
        v = v * gate.repeat_interleave(n, 0)
    if cfg['n1']:
        v = _ln_ref(v, g1)
    w = v + res.float()
    if out is not None:
        assert_close(out, w, 8e-3, 8e-3, f'{case}: raw')
    if cfg['n2']:
        if cfg['n2'] == 1:
            y = _ln_ref(w, g2)
        else:
            y = F.normalize(w, dim=-1) * g2 * math.sqrt(Cc)
            if film is not None:
                fr = film.repeat_interleave(n, dim=0)
                y = y * (fr[:, :Cc] + 1) + fr[:, Cc:2 * Cc]
            y = F.silu(y)
        assert_close(out_n, y, 1.2e-2, 1.2e-2, f'{case}: normalised')


@pytest.mark.parametrize('n,Cc,B', [(4096, 128, 3), (64, 1024, 3), (300, 40, 3), (256, 512, 11), (1024, 256, 32), (16, 2048, 9)])
def test_global_context_gate_and_gate_residual(n, Cc, B):
    hid = max(3, Cc // 2)
    x = rnd(B * n, Cc).to(BF16)
    wk, bk = rnd(Cc, scale=0.2, seed=1), 0.1
    w1, b1 = rnd(hid, Cc, scale=1 / math.sqrt(Cc), seed=2), rnd(hid, scale=0.1, seed=3)
    w2, b2 = rnd(Cc, hid, scale=1 / math.sqrt(hid), seed=4), rnd(Cc, scale=0.1, seed=5)
    nchunk = _lib.load().b200_gca_chunks(n, Cc)
    scratch = torch.zeros(B * nchunk * (Cc + 2) + B * Cc + B * hid + B * n, device=DEV)
    gate = torch.zeros(B, Cc, device=DEV)
    _lib.call('b200_gca_gate', x.data_ptr(), Cc, B, n, Cc, wk.data_ptr(), bk, w1.data_ptr(), b1.data_ptr(), hid, w2.data_ptr(), b2.data_ptr(),
              scratch.data_ptr(), nchunk, gate.data_ptr(), stream())
    xf = x.float().view(B, n, Cc)
    att = (xf @ wk + bk).softmax(dim=-1)                                     # GlobalContext.forward (imagen_pytorch.py:965-970)
    pooled = torch.einsum('bn,bnc->bc', att, xf)
    ref = torch.sigmoid(F.silu(pooled @ w1.T + b1) @ w2.T + b2)
    torch.cuda.synchronize()
    assert_close(gate, ref, rtol=2e-3, atol=2e-3, what='gate')
    res = rnd(B * n, Cc, seed=7).to(BF16)
    out = torch.zeros_like(x)
    _lib.call('b200_gate_residual', x.data_ptr(), Cc, gate.data_ptr(), res.data_ptr(), Cc, out.data_ptr(), Cc, B * n, Cc, n, stream())
    torch.cuda.synchronize()
    assert_close(out, x.float() * gate.repeat_interleave(n, 0) + res.float(), rtol=8e-3, atol=4e-3, what='gate*x+res')


def test_layout_gathers():
    B, H, W = 2, 12, 12
    img0, img1 = rnd(B, 3, H, W), rnd(B, 3, H, W, seed=1)
    ks, Cin = 7, 6
    Kpad = ops.ceil_to(ks * ks * Cin, 64)
    out = torch.zeros(B * H * W, Kpad, dtype=BF16, device=DEV)
    _lib.call('b200_im2col_init', img0.data_ptr(), 3, img1.data_ptr(), 3, B, H, W, ks, out.data_ptr(), Kpad, stream())
    x = torch.cat((img0, img1), dim=1)
    cols = F.unfold(x, ks, padding=ks // 2).view(B, Cin, ks * ks, H * W).permute(0, 3, 2, 1).reshape(B * H * W, ks * ks * Cin)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :ks * ks * Cin], cols.to(BF16)) and out[:, ks * ks * Cin:].abs().max() == 0
    # three sources: x | self_cond | lowres_cond_img (imagen_pytorch.py:1541-1551)
    img2 = rnd(B, 3, H, W, seed=2)
    K3 = ops.ceil_to(ks * ks * 9, 64)
    out3 = torch.zeros(B * H * W, K3, dtype=BF16, device=DEV)
    _lib.call('b200_im2col_init3', img0.data_ptr(), 3, img1.data_ptr(), 3, img2.data_ptr(), 3, B, H, W, ks, out3.data_ptr(), K3, stream())
    x3 = torch.cat((img0, img1, img2), dim=1)
    cols3 = F.unfold(x3, ks, padding=ks // 2).view(B, 9, ks * ks, H * W).permute(0, 3, 2, 1).reshape(B * H * W, ks * ks * 9)
    torch.cuda.synchronize()
    assert torch.equal(out3[:, :ks * ks * 9], cols3.to(BF16)) and out3[:, ks * ks * 9:].abs().max() == 0
    xr = rnd(B, H, W, 16).to(BF16)
    o2 = torch.zeros(B * (H // 2) * (W // 2), 64, dtype=BF16, device=DEV)
    _lib.call('b200_pixel_unshuffle', xr.data_ptr(), 16, B, H, W, 16, o2.data_ptr(), stream())
    ref = xr.view(B, H // 2, 2, W // 2, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64)   # (s1, s2, c) channel order
    torch.cuda.synchronize()
    assert torch.equal(o2, ref)
    o3 = torch.zeros(B * H * W, 8, dtype=BF16, device=DEV)
    _lib.call('b200_nchw_to_rows', img0.data_ptr(), B, 3, H, W, o3.data_ptr(), 8, stream())
    torch.cuda.synchronize()
    assert torch.equal(o3[:, :3], img0.permute(0, 2, 3, 1).reshape(-1, 3).to(BF16)) and o3[:, 3:].abs().max() == 0


@pytest.mark.parametrize('W,ks,chans', [(32, 15, (3,)), (64, 15, (3, 3)), (64, 7, (3, 3, 3, 3)), (96, 3, (5, 3))])
def test_im2col_staged_window(W, ks, chans):
    """Stem patch gather through the shared-memory window (image width a multiple of 32): exact against F.unfold, 1-4 sources."""
    B, H = 3, 20
    imgs = [rnd(B, c, H, W, seed=i) for i, c in enumerate(chans)]
    Cin = sum(chans)
    Kpad = ops.ceil_to(ks * ks * Cin, 64)
    out = torch.full((B * H * W, Kpad), 7.0, dtype=BF16, device=DEV)
    ptrs = []
    for i in range(4):
        ptrs += [imgs[i].data_ptr(), chans[i]] if i < len(imgs) else [None, 0]
    _lib.call('b200_im2col_init4', *ptrs, B, H, W, ks, out.data_ptr(), Kpad, stream())
    x = torch.cat(imgs, dim=1)
    cols = F.unfold(x, ks, padding=ks // 2).view(B, Cin, ks * ks, H * W).permute(0, 3, 2, 1).reshape(B * H * W, ks * ks * Cin)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :ks * ks * Cin], cols.to(BF16)) and out[:, ks * ks * Cin:].abs().max() == 0


def test_time_conditioning_plumbing():
    R, D, S = 4, 96, 5
    table, th = rnd(S, D), rnd(R, D, seed=1)
    slots = torch.tensor([3, 3, 0, 4], dtype=torch.int32, device=DEV)
    out = torch.zeros(R, D, dtype=BF16, device=DEV)
    _lib.call('b200_make_time_cond', table.data_ptr(), th.data_ptr(), slots.data_ptr(), R, D, out.data_ptr(), stream())
    torch.cuda.synchronize()
    assert_close(out, F.silu(table[slots.long()] + th), rtol=8e-3, atol=2e-3)
    tb = rnd(S, 2, 64).to(BF16)
    dst = torch.zeros(R, 10, 64, dtype=BF16, device=DEV)
    jobs = (TimeRowJob * 1)(TimeRowJob(tb.data_ptr(), dst.data_ptr(), 10 * 64, 2, 64))
    jd = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(DEV)
    _lib.call('b200_update_time_rows', jd.data_ptr(), 1, slots.data_ptr(), R, 128, stream())
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :2], tb[slots.long()]) and dst[:, 2:].abs().max() == 0


# ------------------------------------------------------------------------------------------------ fp32 conditioning head

def test_f32_conditioning_kernels():
    M, K, N = 37, 129, 70
    x, Wt, b, res = rnd(M, K), rnd(N, K, scale=0.1, seed=1), rnd(N, seed=2), rnd(M, N, seed=3)
    y = torch.zeros(M, N, device=DEV)
    _lib.call('b200_linear_f32', x.data_ptr(), K, Wt.data_ptr(), b.data_ptr(), _lib.ACT_SILU, _lib.ACT_GELU, res.data_ptr(), N, y.data_ptr(), N, M, N, K, stream())
    torch.cuda.synchronize()
    assert_close(y, F.gelu(F.linear(F.silu(x), Wt, b)) + res, rtol=1e-4, atol=1e-4, what='linear_f32')
    g, beta = rnd(K, seed=4), rnd(K, seed=5)
    y2 = torch.zeros(M, K, device=DEV)
    _lib.call('b200_layernorm_f32', x.data_ptr(), K, g.data_ptr(), beta.data_ptr(), 1e-5, y2.data_ptr(), K, M, K, stream())
    torch.cuda.synchronize()
    assert_close(y2, F.layer_norm(x, (K,), g, beta, 1e-5), rtol=1e-4, atol=1e-4, what='layernorm_f32')
    t, w = torch.tensor([-33.891, 8.769, 0.7093, -1.78], device=DEV), rnd(8, seed=6)
    e = torch.zeros(4, 17, device=DEV)
    _lib.call('b200_sinu_pos_emb', t.data_ptr(), w.data_ptr(), 4, 8, e.data_ptr(), stream())
    fr = t[:, None] * w[None, :] * 2 * math.pi                            # LearnedSinusoidalPosEmb (imagen_pytorch.py:664-669)
    torch.cuda.synchronize()
    assert_close(e, torch.cat((t[:, None], fr.sin(), fr.cos()), dim=-1), rtol=0, atol=2e-5, what='sinu (hundreds of radians)')
    B, H, nq, nk = 2, 8, 36, 60
    q, kv = rnd(B, nq, H * 64), rnd(B, nk, 2 * H * 64, seed=7)
    qs, ks = rnd(64, seed=8).abs() + 0.5, rnd(64, seed=9).abs() + 0.5
    o = torch.zeros(B, nq, H * 64, device=DEV)
    _lib.call('b200_attn_f32', q.data_ptr(), H * 64, kv.data_ptr(), kv.data_ptr() + 4 * H * 64, 2 * H * 64, qs.data_ptr(), ks.data_ptr(), o.data_ptr(),
              H * 64, B, H, nq, nk, stream())
    qh = F.normalize(q.view(B, nq, H, 64), dim=-1) * qs
    kh = F.normalize(kv[..., :H * 64].reshape(B, nk, H, 64), dim=-1) * ks
    vh = kv[..., H * 64:].reshape(B, nk, H, 64)
    ref = torch.einsum('bhij,bjhd->bihd', (torch.einsum('bihd,bjhd->bhij', qh, kh) * 8).softmax(-1), vh).reshape(B, nq, H * 64)
    torch.cuda.synchronize()
    assert_close(o, ref, rtol=1e-4, atol=1e-4, what='attn_f32')
    dst = torch.zeros(B, 50, H * 64, dtype=BF16, device=DEV)
    kv2 = rnd(B * 10, 2 * H * 64, seed=11)
    _lib.call('b200_headnorm_store', kv2.data_ptr(), 2 * H * 64, 0, H, 1, ks.data_ptr(), dst.data_ptr() + 2 * 3 * H * 64, 10, 50 * H * 64, H * 64, 64,
              B * 10, stream())
    torch.cuda.synchronize()
    ref = (F.normalize(kv2[:, :H * 64].view(B, 10, H, 64), dim=-1) * ks).reshape(B, 10, H * 64)
    assert_close(dst[:, 3:13], ref, rtol=8e-3, atol=2e-3, what='headnorm_store')
    assert dst[:, :3].abs().max() == 0 and dst[:, 13:].abs().max() == 0


# ------------------------------------------------------------------------------------------------ sampler steps

def _ddpm_reference(x, pred, noise, coef, cond_scale, B, dynamic=True):
    from oracle.sampler_ref import dynamic_threshold
    sigma, alpha, _, alpha_next, c, noise_std = [coef[i] for i in range(6)]
    e = pred[:B] if pred.shape[0] == B else pred[B:] + (pred[:B] - pred[B:]) * cond_scale
    x0 = (x - sigma * e) / alpha.clamp(min=1e-8)
    x0 = dynamic_threshold(x0) if dynamic else x0.clamp(-1, 1)
    mean = alpha_next * (x * (1 - c) / alpha + c * x0)
    return mean + noise_std * noise


@pytest.mark.parametrize('shape,cfg', [((4, 3, 64, 64), True), ((2, 3, 32, 32), False), ((2, 3, 160, 160), True)])
def test_ddpm_step_is_bit_exact_vs_torch_ops(shape, cfg):
    """Same torch ops as the reference on the same device: the fused step (CFG, x0, exact quantile, posterior) must be bit-identical."""
    from imagen_pytorch_b200.imagen import GaussianDiffusionContinuousTimes, quantile_ranks
    B = shape[0]
    R = 2 * B if cfg else B
    chw = shape[1] * shape[2] * shape[3]
    coefs, _ = GaussianDiffusionContinuousTimes(noise_schedule='cosine', timesteps=10).ddpm_coefficients(DEV)
    x, noise = rnd(*shape), rnd(*shape, seed=1)
    pred = rnd(R, *shape[1:], seed=2)
    x[0, 0, 0, :8] = 3.0                                                 # duplicates around the quantile are legal
    slots = torch.full((R,), 4, dtype=torch.int32, device=DEV)
    q = quantile_ranks(chw, 0.95, DEV)
    expect = _ddpm_reference(x.clone(), pred, noise, coefs[4], 3.0, B)
    _lib.call('b200_ddpm_step', x.data_ptr(), pred.data_ptr(), noise.data_ptr(), coefs.data_ptr(), slots.data_ptr(), R, B, chw, 3.0, 0, 1, q[0], q[1], q[2],
              stream())
    torch.cuda.synchronize()
    assert torch.equal(slots, torch.full_like(slots, 5))
    # torch evaluates sigmoid/exp op by op like the kernel; allow 2 ulp for the fused-multiply differences of torch's own kernels
    assert_close(x, expect, rtol=3e-7, atol=3e-7, what='ddpm step')


def test_quantile_threshold_exact_against_torch_quantile():
    """Drive the step kernel so that its output exposes s = max(1, quantile_0.95(|x0|)) and compare with torch.quantile."""
    from imagen_pytorch_b200.imagen import quantile_ranks
    B, chw = 3, 3 * 64 * 64
    x0 = rnd(B, 3, 64, 64, scale=2.5)
    x0[1] *= 0.1                                                         # s clamps to 1 for this sample
    coefs = torch.tensor([[0., 1., 1., 1., 1., 0., 0., 0.]], device=DEV)  # sigma=0, alpha=1, c=1: x_next = clamp(x0, -s, s)/s
    x = x0.clone()
    zeros = torch.zeros_like(x)
    slots = torch.zeros(B, dtype=torch.int32, device=DEV)
    q = quantile_ranks(chw, 0.95, DEV)
    _lib.call('b200_ddpm_step', x.data_ptr(), zeros.data_ptr(), zeros.data_ptr(), coefs.data_ptr(), slots.data_ptr(), B, B, chw, 1.0, 0, 1, q[0], q[1], q[2],
              stream())
    torch.cuda.synchronize()
    s = torch.quantile(x0.flatten(1).abs(), 0.95, dim=-1).clamp(min=1.).view(-1, 1, 1, 1)
    assert_close(x, x0.clamp(-s, s) / s, rtol=2.5e-7, atol=0, what='threshold (1 ulp: lerp FMA contraction inside ATen)')
    # the self-conditioning variant additionally exports the thresholded x_start (here identical to the step's output)
    x2, xs = x0.clone(), torch.full_like(x0, float('nan'))
    slots.zero_()
    _lib.call('b200_ddpm_step_sc', x2.data_ptr(), zeros.data_ptr(), zeros.data_ptr(), coefs.data_ptr(), slots.data_ptr(), B, B, chw, 1.0, 0, 1, q[0], q[1],
              q[2], xs.data_ptr(), stream())
    torch.cuda.synchronize()
    assert torch.equal(x2, x) and torch.equal(xs, x)
