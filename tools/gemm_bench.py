"""GEMM bottleneck experiments: time representative b200_conv_gemm shapes alone (L2 flushed between launches) under the
B200_IMAGEN_GEMM_DEBUG switches (1 = no epilogue work, 2 = no MMA issue, 4 = no TMA loads) and with / without CTA pairs.
usage: python tools/gemm_bench.py            (spawns one subprocess per configuration)"""
import os
import statistics
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from imagen_pytorch_b200 import ops, _lib
    dev = torch.device('cuda')
    R = 32
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def conv(H, Cin, Cout):
        x = torch.randn(R, H, H, Cin, device=dev).to(torch.bfloat16)
        Wt = torch.randn(Cout, Cin, 3, 3, device=dev) / (3 * Cin ** 0.5)
        segs, mats = ops.conv_segments(Wt, [Cin])
        wp = ops.pack_weight(mats, Cout, dev)
        out = torch.empty(R * H * H, Cout, dtype=torch.bfloat16, device=dev)
        call = ops.GemmCall([(x.data_ptr(), Cin, Cin)], segs, (R, H, H), wp, Cout, out.data_ptr(),
                            bias=ops.padded_bias(torch.zeros(Cout, device=dev), Cout, dev), ldc=Cout)
        return call, 2.0 * R * H * H * Cout * 9 * Cin, (x, wp, out)

    def linear(M, K, N, **epi):
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev) / K ** 0.5
        wp = ops.pack_weight([W], N, dev)
        ps = epi.get('out_mode') == _lib.OUT_PIXEL_SHUFFLE
        out = torch.empty(M * (4 if ps else 1), (N // 4) if ps else N, dtype=torch.bfloat16, device=dev)
        grid = epi.pop('grid', (1, 1, M))
        call = ops.GemmCall([(x.data_ptr(), K, K)], [(0, 0, 0)], grid, wp, N, out.data_ptr(), ldc=out.shape[1], **epi)
        return call, 2.0 * M * K * N, (x, wp, out)

    qs = torch.ones(64, device=dev)
    cases = {
        'conv3x3 64x64 128->128': conv(64, 128, 128),
        'conv3x3 32x32 256->256': conv(32, 256, 256),
        'conv3x3 16x16 512->512': conv(16, 512, 512),
        'conv3x3 8x8 1024->1024': conv(8, 1024, 1024),
        'to_q 131072x128->512 l2norm': linear(131072, 128, 512, l2_cols=512, l2_scale=qs),
        'plain 131072x128->512': linear(131072, 128, 512),
        'ff1 131072x128->256 gelu': linear(131072, 128, 256, act=_lib.ACT_GELU),
        'to_out 131072x512->128': linear(131072, 512, 128),
        'pixshuf 32x32 256->512': linear(32768, 256, 512, out_mode=_lib.OUT_PIXEL_SHUFFLE, ps_C=128, grid=(R, 32, 32),
                                         bias=ops.padded_bias(torch.zeros(512, device=dev), 512, dev), act=_lib.ACT_SILU),
    }
    st = torch.cuda.current_stream(dev)
    tag = (f"pair={os.environ.get('B200_IMAGEN_GEMM_PAIR', '0')} tma_store={os.environ.get('B200_IMAGEN_GEMM_TMA_STORE', '1')} "
           f"epi16={os.environ.get('B200_IMAGEN_GEMM_EPI16', '1')} debug={os.environ.get('B200_IMAGEN_GEMM_DEBUG', '0')}")
    only = os.environ.get('GEMM_BENCH_ONLY')
    for name, (call, flops, keep) in cases.items():
        if only and only not in name:
            continue
        for _ in range(3):
            call(st.cuda_stream)
        times = []
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            call(st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize(dev)
            times.append(e0.elapsed_time(e1))
        us = statistics.median(times) * 1e3
        print(f'{tag}  {name:30s} {us:7.1f} us  {flops / us / 1e6:7.0f} TFLOP/s', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for epi16, dbg in (('1', '0'), ('0', '0'), ('1', '6'), ('0', '6')):
            env = dict(os.environ, B200_IMAGEN_GEMM_EPI16=epi16, B200_IMAGEN_GEMM_DEBUG=dbg)
            subprocess.run([sys.executable, __file__, 'child'], env=env, check=False)
