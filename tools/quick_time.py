"""Quick device timing of the fused generator forward (development aid; bench.py is the contract benchmark)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--impl', type=int, default=0)
    ap.add_argument('--passes', type=int, default=3)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--graphs', type=int, default=1)
    ap.add_argument('--layers', action='store_true', help='per-call timing via launch blocking events')
    a = ap.parse_args()
    G = S.random_generator(a.res, seed=0, device='cuda')
    G.engine(passes=a.passes, impl=a.impl, graphs=bool(a.graphs) and not a.layers)
    x, z = S.synthetic_batch(a.batch, a.res, seed=0)
    x, z = x.cuda(), z.cuda()
    for _ in range(2):
        G(x, z, None, noise_mode='random')
    torch.cuda.synchronize()
    if a.layers:
        from shgan_b200 import kernels as K
        names = ['conv_igemm', 'conv_up2', 'fir_nhwc', 'fromrgb', 'torgb_combine', 'dense', 'style_prep', 'shu_fwd', 'nchw_to_planes',
                 'planes_to_nchw', 'planes_add_nchw', 'normalize_2nd_moment']
        acc = {}
        import shgan_b200.engine as E
        for nm in names:
            fn = getattr(K, nm)

            def wrap(*args, _fn=fn, _nm=nm, **kw):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = _fn(*args, **kw)
                e.record()
                e.synchronize()
                key = _nm
                if _nm == 'conv_igemm':
                    srcs, w_hi = args[0], args[1]
                    key = f'conv C{srcs[0].shape[3]}->{w_hi.shape[1]} {args[4]}x{args[5]} taps{len(args[3])} {"raw" if kw.get("raw") is not None else "act"}'
                elif _nm == 'conv_up2':
                    key = f'up2 C{args[0].shape[3]}->{args[1].shape[0] * 64} in {args[0].shape[1]}x{args[0].shape[2]}'
                elif _nm == 'fir_nhwc':
                    key = f'fir {tuple(args[0].shape)}'
                t = acc.setdefault(key, [0.0, 0])
                t[0] += s.elapsed_time(e)
                t[1] += 1
                return r
            setattr(K, nm, wrap)
        G(x, z, None, noise_mode='random')
        tot = sum(v[0] for v in acc.values())
        for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            print(f'{v[0]:9.3f} ms  x{v[1]:3d}  {k}')
        print(f'total {tot:.3f} ms')
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(a.iters):
        s.record()
        G(x, z, None, noise_mode='random')
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ms = min(ts)
    flop = {512: 238.785e9, 256: 180.635e9}.get(a.res, 0) * a.batch
    print(f'res {a.res} batch {a.batch} impl {a.impl} passes {a.passes} graphs {a.graphs}: best {ms:.3f} ms median {sorted(ts)[len(ts)//2]:.3f} ms '
          f'-> {a.batch / ms * 1e3:.1f} img/s, {flop / ms / 1e9:.1f} algorithmic TFLOP/s')


if __name__ == '__main__':
    main()
