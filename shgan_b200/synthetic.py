"""Synthetic inputs and random-init weights for benchmarks and smoke runs (no datasets or checkpoints are
reachable offline).  Inputs follow the eval loop's construction, lib/experiments/shgan_default.py:269-276:
x = cat([mask - 0.5, image * mask]) with a free-form mask (1 = keep), z ~ N(0, 1).
"""
import math

import numpy as np
import torch

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


def freeform_mask(res, rng, hole_range=(0.0, 1.0)):
    """Free-form mask in the spirit of RandomMask / RandomBrush (lib/data_factory/ds_ffhq.py:145-217): random thick
    poly-line strokes plus random rectangles, re-drawn until the hole ratio falls in `hole_range`.  1 = keep."""
    s = res
    for _ in range(50):
        m = np.ones((s, s), np.float32)
        for _ in range(int(rng.integers(1, 5))):      # rectangles
            w, h = int(rng.integers(s // 8, s // 2)), int(rng.integers(s // 8, s // 2))
            x0, y0 = int(rng.integers(0, s - w)), int(rng.integers(0, s - h))
            m[y0:y0 + h, x0:x0 + w] = 0
        for _ in range(int(rng.integers(1, 4))):      # brush strokes
            px, py = float(rng.integers(0, s)), float(rng.integers(0, s))
            width = int(rng.integers(max(2, s // 40), max(3, s // 10)))
            for _ in range(int(rng.integers(2, 8))):
                ang, ln = rng.uniform(0, 2 * math.pi), rng.uniform(s / 16, s / 3)
                qx = float(np.clip(px + ln * math.cos(ang), 0, s - 1))
                qy = float(np.clip(py + ln * math.sin(ang), 0, s - 1))
                for t in np.linspace(0, 1, int(max(abs(qx - px), abs(qy - py))) + 1):
                    cx, cy = int(px + (qx - px) * t), int(py + (qy - py) * t)
                    m[max(cy - width // 2, 0):cy + width // 2 + 1, max(cx - width // 2, 0):cx + width // 2 + 1] = 0
                px, py = qx, qy
        hole = 1.0 - float(m.mean())
        if hole_range[0] <= hole <= hole_range[1]:
            return m
    return m


def synthetic_batch(batch, res, seed=0, z_dim=512):
    """-> (x [B,4,R,R] float32, z [B,z_dim] float32) CPU tensors."""
    rng = np.random.default_rng(seed)
    img = np.clip(rng.standard_normal((batch, 3, res, res)), -1, 1).astype(np.float32)
    mask = np.stack([freeform_mask(res, rng) for _ in range(batch)])[:, None]
    x = np.concatenate([mask - 0.5, img * mask], axis=1).astype(np.float32)
    z = rng.standard_normal((batch, z_dim)).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(z)


def generator_cfg(res, ch_base=32768, ch_max=512):
    """Model config of `shgan_g256` / `shgan_g512` (configs/model/{shgan,comodgan,stylegan}.yaml) as plain dicts."""
    num_ws = 2 * int(math.log2(res)) - 2
    m = dict(type='comodgan_mapping', args=dict(z_dim=512, c_dim=0, w_dim=512, num_ws=num_ws, num_layers=8, embed_features=None,
                                                layer_features=None, activation=ACT, lr_multiplier=0.01, w_avg_beta=0.995))
    e = dict(type='shgan_encoder', args=dict(resolution=res, ic_n=4, oc_n=1024, ch_base=ch_base, ch_max=ch_max,
                                             use_fp16_before_res=None, resample_filter=[1, 3, 3, 1], activation=ACT,
                                             mbstd_group_size=0, mbstd_c_n=0, c_dim=None, cmap_dim=None, use_dropout=True,
                                             has_extra_final_layer=False, shu_channels=32, shu_df_freedom=[2, 3],
                                             shu_df_type='piecewise_linear', shu_input_res=64, shu_lowest_res=4,
                                             shu_tail_sigma_mult=3, shu_gaussian_at_input_res=False))
    s = dict(type='comodgan_synthesis', args=dict(w_dim=512, w0_dim=1024, resolution=res, rgb_n=3, ch_base=ch_base, ch_max=ch_max,
                                                  use_fp16_after_res=None, resample_filter=[1, 3, 3, 1], activation=ACT))
    return dict(type='comodgan_generator', args=dict(mapping=m, encoder=e, synthesis=s))


def random_generator(res, seed=0, device='cuda', ch_base=32768, ch_max=512):
    """Random-init generator of the reference architecture (constructor statistics), with the terms a fresh init
    zeroes out (biases, noise strengths) randomised so that every kernel term does real work."""
    from .model_zoo import get_model
    torch.manual_seed(seed)
    G = get_model()(generator_cfg(res, ch_base, ch_max))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith('noise_strength'):
                p.copy_(torch.randn([], generator=g) * 0.1)
            elif name.endswith('.bias') and not name.endswith('affine.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return G.eval().requires_grad_(False).to(device)
