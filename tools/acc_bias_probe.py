"""Calibration of shgan_conv_desc::acc_comp (the compensation of the tensor core's truncating fp32 accumulate).

(1) one wide layer per kernel against an fp64 convolution: slope of the least-squares fit y ~ (1 + s) * ref, max-abs error;
(2) the batch-16 512^2 generator against the reference fixture (tests/golden/gen512_b16.npz) for several acc_comp values.
    python tools/acc_bias_probe.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from shgan_b200 import kernels as K, packing as P  # noqa: E402
from oracle import shgan_oracle as O  # noqa: E402
import helpers as H  # noqa: E402
from golden.make_golden import bench_subsample  # noqa: E402

dev = 'cuda'
COMPS = [-1.0, 1.5e-8, 2.15e-8, 3.0e-8, 4.0e-8]


def layer_probe():
    g = torch.Generator().manual_seed(0)
    for (ci, co, hw, n, label) in [(512, 512, 32, 4, '512->512 @32'), (128, 128, 64, 4, '128->128 @64'), (64, 64, 128, 2, '64->64 @128')]:
        for dist in ('randn', 'relu'):
            x = torch.randn(n, ci, hw, hw, generator=g)
            if dist == 'relu':
                x = torch.nn.functional.leaky_relu(x, 0.2) * 1.4
            w = torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** 0.5)
            ref = torch.nn.functional.conv2d(x.double().to(dev), w.double().to(dev), padding=1)
            xp = K.nchw_to_planes(x.to(dev))
            wh, wl = P.pack_conv_weight(w.to(dev))
            for impl in (0, 2, 3, 4):
                for comp in (-1.0, 2.15e-8):
                    y = torch.empty((n, hw, hw, co), device=dev)
                    K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), hw, hw, epi=K.make_epilogue(out_f32=y), impl=impl, acc_comp=comp)
                    yy = y.permute(0, 3, 1, 2).double()
                    slope = float((yy * ref).sum() / (ref * ref).sum() - 1)
                    err = float((yy - ref).abs().max())
                    print(f'{label:14s} {dist:5s} impl {impl} comp {comp:9.2e}: slope {slope:+.3e}  max-abs {err:.3e} (|ref|max {float(ref.abs().max()):.2f})')


def generator_probe():
    name, res, batch, seed = 'gen512_b16', 512, 16, 14
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    G = H.build_generator(res, O.synthetic_state_dict(res, seed=seed), device=dev)
    eng = G.engine(passes=3, impl=0)
    x, z = O.synthetic_inputs(batch, res, seed=seed)
    xd, zd = torch.from_numpy(x).to(dev), torch.from_numpy(z).to(dev)
    for comp in COMPS:
        K.ACC_COMP = comp
        eng._graphs = {}
        img, _ = G.forward_composite(xd, zd, noise_mode='const')
        img = img.cpu().numpy()
        err = np.abs(bench_subsample(img).astype(np.float64) - gold['img_sub']).max()
        st = gold['stats']
        bias = np.mean([(np.abs(img[n]).astype(np.float64).sum() - st[n, 3]) / st[n, 3] for n in range(batch)])
        print(f'generator 512 b16 acc_comp {comp:9.2e}: max-abs err {err:.3e}  mean relative |img| bias {bias:+.3e}')


if __name__ == '__main__':
    layer_probe()
    generator_probe()
