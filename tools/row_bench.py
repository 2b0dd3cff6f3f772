"""Time the row-wise norm kernels alone (cold L2: 8 rotating buffer sets) for each B200_IMAGEN_ROW_VPT target.
usage: python tools/row_bench.py            (spawns one subprocess per target)"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(131072, 128), (131072, 256), (32768, 256), (32768, 512), (8192, 512), (8192, 1024), (2048, 1024)]


def child():
    import torch
    from imagen_pytorch_b200 import _lib
    dev = torch.device('cuda')
    for M, C in SHAPES:
        nb = max(2, int(300e6 // (M * C * 4)) + 1)
        xs = [torch.randn(M, C, device=dev).bfloat16() for _ in range(nb)]
        outs = [torch.empty_like(x) for x in xs]
        g = torch.randn(C, device=dev)
        film = torch.randn(32, 2 * C, device=dev)
        n = M // 32
        res = {}
        for name in ('rms', 'ln'):
            def run(i):
                st = torch.cuda.current_stream().cuda_stream
                if name == 'rms':
                    sa = (_lib.Src * 1)(_lib.Src(xs[i].data_ptr(), C, C))
                    _lib.call('b200_rmsnorm_film_silu', sa, 1, 1.0, g.data_ptr(), film.data_ptr(), 2 * C, n, outs[i].data_ptr(), C, M, st)
                else:
                    _lib.call('b200_layernorm', xs[i].data_ptr(), C, g.data_ptr(), None, 1e-5, xs[(i + 1) % nb].data_ptr(), C, outs[i].data_ptr(), C, M, C, st)
            for i in range(nb):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 * nb
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for i in range(reps):
                    run(i % nb)
            gr.replay()
            torch.cuda.synchronize()
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            byts = M * C * 2 * (2 if name == 'rms' else 3)
            res[name] = (us, byts / us / 1e3)
        print(f'vpt {os.environ.get("B200_IMAGEN_ROW_VPT", "-")}  M {M:7d} C {C:5d}  rmsnorm {res["rms"][0]:7.1f} us {res["rms"][1]:6.0f} GB/s   '
              f'layernorm+res {res["ln"][0]:7.1f} us {res["ln"][1]:6.0f} GB/s', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for t in (1, 2, 4, 8):
            env = dict(os.environ, B200_IMAGEN_ROW_VPT=str(t))
            subprocess.run([sys.executable, __file__, 'child'], env=env, check=False)
