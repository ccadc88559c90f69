"""Times shgan_conv_up2 alone at the generator's layer sizes (development aid; also the ncu target for that kernel)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P  # noqa: E402

dev = 'cuda'
LAYERS = {'b512': (16, 128, 64, 256), 'b256': (16, 256, 128, 128), 'b128': (16, 512, 256, 64), 'b64': (16, 512, 512, 32)}
LAYERS['b32'] = (16, 512, 512, 16)
which = [a for a in sys.argv[1:] if a in LAYERS] or list(LAYERS)
NARROW = 'narrow' in sys.argv       # force the 8-warp epilogue instance
CLUSTER = True if 'cluster' in sys.argv else (False if 'nocluster' in sys.argv else None)     # force / forbid the weight-sharing CTA pairs
NO_SKIP = 'noskip' in sys.argv      # experiment: epilogue without the skip planes / the noise (how much do their loads cost?)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in which:
    n, ci, co, h = LAYERS[name]
    g = torch.Generator().manual_seed(0)
    x = K.Planes.empty(n, h, h, ci, dev)
    x.hi.normal_(); x.lo.normal_(0, 1e-3)
    w = torch.randn(co, ci, 3, 3, generator=g).to(dev) / (3 * ci ** 0.5)
    uh, ul = P.pack_up2_weight(w)
    out = K.Planes.empty(n, 2 * h, 2 * h, co, dev)
    skip = K.Planes.empty(n, 2 * h, 2 * h, co, dev)
    skip.hi.normal_()
    dc = torch.rand(n, co, device=dev) + 0.5
    ns = torch.rand(n, co, device=dev) + 0.5
    bias = torch.randn(co, device=dev)
    nz = torch.randn(n, 1, 2 * h, 2 * h, device=dev)
    st = torch.tensor(0.1, device=dev)
    epi = K.make_epilogue(dcoef=dc, noise=None if NO_SKIP else nz, noise_sn=4 * h * h, noise_strength=st, bias=bias, act=True, act_gain=2 ** 0.5,
                          act_clamp=256.0, skip=None if NO_SKIP else skip, next_scale=ns, out=out)
    fy = fx = [0.125, 0.375, 0.375, 0.125]
    run = lambda: K.conv_up2(x, uh, ul, fx, fy, 4.0, epi, narrow=NARROW, cluster=CLUSTER)
    for _ in range(3):
        run()
    ts = []
    for _ in range(7):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * n * h * h * 9 * ci * co
    byts = 4.0 * n * (h * h * ci + 2 * 4 * h * h * co)
    print(f'{name}: {ms:.3f} ms  {flops / ms / 1e9:.0f} algorithmic TFLOP/s  {byts / ms / 1e6:.0f} GB/s (in + skip + out)')
