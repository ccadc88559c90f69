"""Device-time micro-benchmarks of the memory-bound kernels (development aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P

def timeit(fn, iters=10, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e3   # us

dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
which = sys.argv[1:] or ['dense', 'fir', 'fromrgb', 'upfirdn']
if 'dense' in which:
    for (B, I, O, I0) in [(16, 512, 512, 512), (16, 1536, 512, 512), (16, 1536, 64, 512), (16, 8192, 1024, 8192), (16, 1024, 8192, 1024), (16, 1024, 8960, 512), (4, 1536, 512, 512)]:
        x0 = torch.randn(B, I0, device=dev); x1 = torch.randn(B, max(I - I0, 4), device=dev) if I > I0 else None
        w = torch.randn(O, I, device=dev); b = torch.randn(O, device=dev); out = torch.empty(B, O, device=dev)
        us = timeit(lambda: K.dense(x0, w, b, out, 0.1, 1.0, True, x1=x1), flush=flush)
        print(f'dense B{B} I{I} O{O}: {us:8.1f} us   weights {w.numel()*4/1e6:.1f} MB -> {w.numel()*4/us/1e3:.0f} GB/s')
if 'fir' in which:
    f = P.setup_filter([1, 3, 3, 1]).to(dev)
    for (N, H, C) in [(16, 512, 64), (16, 256, 128), (16, 64, 512), (16, 32, 512), (16, 16, 512), (16, 8, 512)]:
        src = K.Planes.empty(N, H, H, C, dev); src.hi.normal_(); src.lo.normal_(0, 1e-3)
        ph = (H + 2) // 2
        par = K.Planes.empty(4 * N, ph, ph, C, dev)
        epi = K.make_epilogue(out=par)
        byts = N * H * H * C * 4 + N * (H + 1) ** 2 * C * 4
        for tp in (False, True):
            us = timeit(lambda: K.fir_nhwc(src, f, 1.0, (2, 2, 2, 2), epi, parity_split=True, rank1=True, two_phase=tp), flush=flush)
            print(f'fir<planes,parity,{"two-phase" if tp else "walk"}> N{N} {H}x{H}x{C}: {us:8.1f} us  {byts/us/1e3:.0f} GB/s')
        z = torch.randn(N, H + 1, H + 1, C, device=dev)
        out = K.Planes.empty(N, H, H, C, dev); skip = K.Planes.empty(N, H, H, C, dev)
        dc = torch.rand(N, C, device=dev); bias = torch.randn(C, device=dev); ns = torch.rand(N, C, device=dev)
        nz = torch.randn(N, 1, H, H, device=dev); st = torch.tensor(0.1, device=dev)
        epi2 = K.make_epilogue(dcoef=dc, noise=nz, noise_sn=H * H, noise_strength=st, bias=bias, act=True, act_gain=1.414, act_clamp=256.0,
                               skip=skip, next_scale=ns, out=out)
        us = timeit(lambda: K.fir_nhwc(z, f, 4.0, (1, 1, 1, 1), epi2), flush=flush)
        byts = N * (H + 1) ** 2 * C * 4 + 2 * N * H * H * C * 4
        print(f'fir<f32,epilogue>   N{N} {H}x{H}x{C}: {us:8.1f} us  {byts/us/1e3:.0f} GB/s')
if 'fromrgb' in which:
    x = torch.randn(16, 4, 512, 512, device=dev); w = torch.randn(64, 4, device=dev); b = torch.randn(64, device=dev)
    out = K.Planes.empty(16, 512, 512, 64, dev)
    us = timeit(lambda: K.fromrgb(x, w, b, 0.5, 0.2, 1.414, 256.0, out), flush=flush)
    print(f'fromrgb 16x512x512 4->64: {us:8.1f} us  {(x.numel()*4 + out.hi.numel()*4)/us/1e3:.0f} GB/s')
if 'upfirdn' in which:
    from shgan_b200 import ops
    f = ops.setup_filter([1, 3, 3, 1], device=torch.device(dev))
    for shape, kw in [((16, 64, 512, 512), dict(padding=[2, 2, 2, 2])), ((16, 64, 513, 513), dict(padding=[1, 1, 1, 1], gain=4)),
                      ((16, 3, 256, 256), dict(up=2, padding=[2, 1, 2, 1], gain=4)), ((16, 64, 512, 512), dict(down=2, padding=[1, 1, 1, 1]))]:
        x = torch.randn(shape, device=dev)
        y = ops.upfirdn2d(x, f, **kw)
        us = timeit(lambda: ops.upfirdn2d(x, f, **kw), flush=flush)
        print(f'upfirdn2d {shape} {kw}: {us:8.1f} us  {(x.numel() + y.numel()) * 4 / us / 1e3:.0f} GB/s')
