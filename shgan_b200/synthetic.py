"""Synthetic inputs and random-init weights for benchmarks and smoke runs (no datasets or checkpoints are
reachable offline).  Inputs follow the eval loop's construction, lib/experiments/shgan_default.py:269-276:
x = cat([mask - 0.5, image * mask]) with a free-form mask (1 = keep), z ~ N(0, 1).
"""
import math

import numpy as np
import torch

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


def _brush_strokes(rs, max_tries, s, vertex_range=(4, 18), mean_angle=2 * math.pi / 5, angle_range=2 * math.pi / 15,
                   width_range=(12, 48)):
    """`RandomBrush` (lib/data_factory/ds_ffhq.py:145-196) restated on an explicit legacy RandomState: poly-line strokes
    with round joints, drawn with PIL like the reference (the rasteriser is part of the result).  The draw ORDER of the
    random numbers is the reference's, so RandomState(seed) reproduces np.random.seed(seed) bit for bit.  Returns a
    uint8 [s,s] array, 1 = brushed (hole)."""
    from PIL import Image, ImageDraw
    mean_radius = math.hypot(s, s) / 8
    canvas = Image.new('L', (s, s), 0)
    for _ in range(rs.randint(max_tries)):
        n_vertex = rs.randint(*vertex_range)
        lo = mean_angle - rs.uniform(0, angle_range)
        hi = mean_angle + rs.uniform(0, angle_range)
        turns = [(2 * math.pi - rs.uniform(lo, hi)) if k % 2 == 0 else rs.uniform(lo, hi) for k in range(n_vertex)]
        side_a, side_b = canvas.size
        pts = [(int(rs.randint(0, side_b)), int(rs.randint(0, side_a)))]
        for ang in turns:
            step = np.clip(rs.normal(loc=mean_radius, scale=mean_radius // 2), 0, 2 * mean_radius)
            px = np.clip(pts[-1][0] + step * math.cos(ang), 0, side_b)
            py = np.clip(pts[-1][1] + step * math.sin(ang), 0, side_a)
            pts.append((int(px), int(py)))
        pen = ImageDraw.Draw(canvas)
        width = int(rs.uniform(*width_range))
        pen.line(pts, fill=1, width=width)
        for vx, vy in pts:
            pen.ellipse((vx - width // 2, vy - width // 2, vx + width // 2, vy + width // 2), fill=1)
        rs.random_sample()        # the reference draws two coins here whose transposes are discarded (ds_ffhq.py:185-188)
        rs.random_sample()
    out = np.asarray(canvas, np.uint8)
    if rs.random_sample() > 0.5:
        out = np.flip(out, 0)
    if rs.random_sample() > 0.5:
        out = np.flip(out, 1)
    return out


def random_mask(s, rs, hole_range=(0.0, 1.0)):
    """`RandomMask(s, hole_range)` (lib/data_factory/ds_ffhq.py:198-217): random rectangles (up to 10*coef of side < s/2,
    up to 5*coef of side < s) AND-ed with the complement of the brush strokes; re-drawn until the hole ratio lies strictly
    inside `hole_range`.  rs: np.random.RandomState (RandomState(k) == the reference under np.random.seed(k)).
    Returns float32 [1,s,s], 1 = keep."""
    coef = min(hole_range[0] + hole_range[1], 1.0)
    while True:
        keep = np.ones((s, s), np.uint8)
        for tries, max_size in ((int(10 * coef), s // 2), (int(5 * coef), s)):
            for _ in range(rs.randint(tries)):
                w, h = rs.randint(max_size), rs.randint(max_size)
                x = rs.randint(-(w // 2), s - w + w // 2)
                y = rs.randint(-(h // 2), s - h + h // 2)
                keep[max(y, 0):min(y + h, s), max(x, 0):min(x + w, s)] = 0
        keep = np.logical_and(keep, 1 - _brush_strokes(rs, int(20 * coef), s))
        hole = 1 - np.mean(keep)
        if hole_range is not None and (hole <= hole_range[0] or hole >= hole_range[1]):
            continue
        return keep[np.newaxis].astype(np.float32)


def freeform_mask(res, rng=None, hole_range=(0.0, 1.0)):
    """Free-form mask [res,res] float32 (1 = keep) = the reference's RandomMask; `rng` is a np.random.RandomState (or an
    int seed)."""
    rs = rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(0 if rng is None else int(rng))
    return random_mask(res, rs, hole_range)[0]


def synthetic_batch(batch, res, seed=0, z_dim=512):
    """-> (x [B,4,R,R] float32, z [B,z_dim] float32) CPU tensors."""
    rng = np.random.default_rng(seed)
    img = np.clip(rng.standard_normal((batch, 3, res, res)), -1, 1).astype(np.float32)
    rs = np.random.RandomState(seed)             # masks: RandomMask(res, [0, 1]) as under np.random.seed(seed)
    mask = np.stack([freeform_mask(res, rs) for _ in range(batch)])[:, None]
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
