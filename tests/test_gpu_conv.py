"""GPU parity tests of the convolution hot path (`-m gpu`).  Test names carry `simt` (fp32 FMA cross-check
kernel, desc.impl=1) or `tc` (tcgen05 tensor-core kernels, the product path: `tc` = per-layer choice as shipped,
`tc_tap` / `tc_halo` = the per-tap / the halo-tile kernel forced for every layer) so the two can be run in separate
processes: `pytest -m gpu -k "not tc"` then `pytest -m gpu -k tc`.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import shgan_oracle as O  # noqa: E402  (checker only)
from golden.make_golden import CONV_CASES, MODCONV_CASES, modconv_inputs, rng, GENERATOR_CASES, DISCRIMINATOR_CASES  # noqa: E402
import helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'
IMPLS = [pytest.param(1, id='simt'), pytest.param(0, id='tc'), pytest.param(2, id='tc_tap'), pytest.param(3, id='tc_halo'), pytest.param(4, id='tc_pair')]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


# ------------------------------------------------------------------------------- raw igemm vs oracle conv
def _igemm_plain(x, w, impl, passes=3, block_n=0):
    """x NCHW, w [Co,Ci,3,3] -> NCHW via the planes kernels (3x3, stride 1, pad 1)."""
    from shgan_b200 import kernels as K, packing as P
    xp = K.nchw_to_planes(t(x))
    wh, wl = P.pack_conv_weight(t(w))
    n, _, h, wd = x.shape
    y = torch.empty((n, h, wd, w.shape[0]), device=DEV)
    K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), h, wd, epi=K.make_epilogue(out_f32=y), passes=passes, impl=impl, block_n=block_n)
    return K.nhwc_to_nchw_f32(y).cpu().numpy()


SHAPES = [
    # n, ci, co, h, w
    (3, 64, 64, 16, 16),     # one tile per image pair
    (2, 128, 64, 24, 40),    # 2 K slabs, ragged tiles in x and y
    (5, 64, 128, 4, 4),      # TN = 8 images per tile, ragged batch, BN = 128
    (3, 64, 256, 8, 8),      # TN = 2, BN = 256
    (1, 192, 64, 5, 7),      # odd sizes -> pow2 tile larger than the image
    (2, 64, 512, 9, 33),     # 2 N blocks of 256
    (1, 64, 64, 70, 45),     # halo kernel: several tiles in x and y with ragged edges
    (2, 128, 128, 33, 64),   # halo kernel: BN = 128, 2 K slabs
    (1, 256, 192, 40, 40),   # 4 K slabs (chunked accumulation across slabs), Co = 3 x 64
    (1, 128, 256, 40, 24),   # two-SM kernel: 8 M tiles -> 4 CTA pairs, 2 K slabs
    (3, 64, 512, 20, 12),    # two-SM kernel: odd number of M tiles (the last pair has an idle half), 2 N blocks
    (2, 128, 128, 24, 24),   # two-SM kernel, BN = 128 (four TMEM accumulator buffers)
    (1, 64, 384, 17, 30),    # two-SM kernel, BN = 128, 3 N blocks, ragged tiles
]


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('shape', SHAPES, ids=[str(s) for s in SHAPES])
def test_igemm_plain(shape, impl):
    n, ci, co, h, w = shape
    g = np.random.default_rng(abs(hash(shape)) % 1000)
    x = g.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = g.standard_normal((co, ci, 3, 3)).astype(np.float32)
    y = _igemm_plain(x, wt, impl)
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), padding=1)
    assert relerr(y, ref) <= 3e-6, relerr(y, ref)


@pytest.mark.parametrize('shape', [(16, 512, 512, 4, 4), (16, 256, 512, 8, 8), (5, 128, 128, 4, 4), (3, 64, 256, 5, 7), (2, 64, 128, 24, 24)],
                         ids=str)
def test_igemm_split_k(shape):
    """ACT mode with a split-K scratch (4x4 / 8x8 layers: several CTAs per output tile, fp32 partials, reduce + fused epilogue
    kernel) against the fp64 convolution with bias + lrelu + fused torgb partial sums, and bit-identical across two runs
    (fixed summation order).  The last shape has enough tiles: the library falls back to the un-split launch."""
    from shgan_b200 import kernels as K, packing as P
    n, ci, co, h, w = shape
    g = np.random.default_rng(sum(shape))
    x = g.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (g.standard_normal((co, ci, 3, 3)) / np.sqrt(ci * 9)).astype(np.float32)
    bias = g.standard_normal(co).astype(np.float32)
    rgb_w = g.standard_normal((3, co)).astype(np.float32)
    rgb_style = g.standard_normal((n, co)).astype(np.float32)
    xp = K.nchw_to_planes(t(x))
    wh, wl = P.pack_conv_weight(t(wt))
    scratch = torch.empty(4 << 20, device=DEV)
    outs = []
    for _ in range(2):
        y = torch.empty((n, h, w, co), device=DEV)
        part = torch.zeros((n, h, w, co // 32, 4), device=DEV)
        epi = K.make_epilogue(bias=t(bias), act=True, act_alpha=0.2, act_gain=float(np.sqrt(2)), act_clamp=256.0, out_f32=y,
                              rgb_w=t(rgb_w), rgb_style=t(rgb_style), rgb_out=part)
        K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), h, w, epi=epi, splitk=scratch)
        outs.append((K.nhwc_to_nchw_f32(y).cpu().numpy(), part.cpu().numpy()))
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), padding=1) + bias[None, :, None, None]
    ref = np.where(ref >= 0, ref, 0.2 * ref) * np.sqrt(2)
    assert relerr(outs[0][0], ref) <= 3e-6
    rgb = np.einsum('nchw,jc,nc->njhw', ref, rgb_w.astype(np.float64), rgb_style.astype(np.float64))
    assert relerr(outs[0][1][..., :3].sum(3).transpose(0, 3, 1, 2), rgb) <= 1e-5
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_igemm_tc_single_pass_and_block_n():
    g = np.random.default_rng(11)
    x = g.standard_normal((2, 64, 16, 16)).astype(np.float32)
    wt = g.standard_normal((256, 64, 3, 3)).astype(np.float32)
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), padding=1)
    y1 = _igemm_plain(x, wt, 0, passes=1)
    assert 1e-5 < relerr(y1, ref) <= 3e-3      # fp16-rounded operands: visibly worse than the split path, still sane
    for bn in (64, 128, 256):
        assert relerr(_igemm_plain(x, wt, 0, block_n=bn), ref) <= 3e-6
    y1 = _igemm_plain(x, wt, 3, passes=1)
    assert 1e-5 < relerr(y1, ref) <= 3e-3
    for bn in (64, 128):
        assert relerr(_igemm_plain(x, wt, 3, block_n=bn), ref) <= 3e-6
    assert 1e-5 < relerr(_igemm_plain(x, wt, 4, passes=1), ref) <= 3e-3     # two-SM kernel, single pass


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_resample_golden(case, impl, golden):
    from shgan_b200 import ops
    name, c = case
    i = [k[0] for k in CONV_CASES].index(name)
    g = rng(200 + i)
    x = g.standard_normal((2, c['ci'], c['hw'], c['hw'])).astype(np.float32)
    w = g.standard_normal((c['co'], c['ci'], c['k'], c['k'])).astype(np.float32)
    f = O.setup_filter([1, 3, 3, 1]) if (c['up'] > 1 or c['down'] > 1) else None
    y = ops.conv2d_resample(t(x), t(w), f=None if f is None else t(f), up=c['up'], down=c['down'], padding=c['k'] // 2,
                            flip_weight=c['flip_weight'], impl=impl).cpu().numpy()
    ref = golden('conv2d_resample')[name]
    assert y.shape == ref.shape
    assert relerr(y, ref) <= 5e-6, relerr(y, ref)


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('case', MODCONV_CASES, ids=[c[0] for c in MODCONV_CASES])
def test_modulated_conv2d_golden(case, impl, golden):
    from shgan_b200 import ops
    name, c = case
    i = [k[0] for k in MODCONV_CASES].index(name)
    x, w, s, nz = modconv_inputs(i, c)
    f = O.setup_filter([1, 3, 3, 1]) if c['up'] > 1 else None
    y = ops.modulated_conv2d(t(x), t(w), t(s), noise=None if nz is None else t(nz), up=c['up'], padding=c['k'] // 2,
                             resample_filter=None if f is None else t(f), demodulate=c['demod'], flip_weight=(c['up'] == 1),
                             impl=impl).cpu().numpy()
    ref = golden('modulated_conv2d')[name]
    assert y.shape == ref.shape
    assert relerr(y, ref) <= 1e-5, relerr(y, ref)


# ------------------------------------------------------------------------------- whole generator vs golden
def _run_generator(name, impl, golden, passes=3):
    _, res, chb, chm, batch, seed = [c for c in GENERATOR_CASES if c[0] == name][0]
    sd = O.synthetic_state_dict(res, seed=seed, ch_base=chb, ch_max=chm)
    G = H.build_generator(res, sd, chb, chm, device=DEV)
    G.engine(passes=passes, impl=impl)
    x, z = O.synthetic_inputs(batch, res, seed=seed)
    g = golden(name)
    img, comp = G.forward_composite(t(x), t(z), noise_mode='const')
    img2 = G(t(x), t(z), torch.zeros(batch, 0, device=DEV), noise_mode='const')
    assert torch.equal(img, img2)                                    # deterministic
    xg, feats = G.encoder(t(x))
    return g, img.cpu().numpy(), comp.cpu().numpy(), xg.cpu().numpy(), {r: v.cpu().numpy() for r, v in feats.items()}


def _check_generator(name, impl, golden, tol):
    g, img, comp, xg, feats = _run_generator(name, impl, golden)
    err = np.abs(img.astype(np.float64) - g['img']).max()
    print(f'{name} impl={impl}: |img|max {np.abs(g["img"]).max():.3f} max-abs err {err:.3e}')
    # intermediate tensors: a few 1e-5 relative (the image tolerance of 1e-3 max-abs is the north-star bar)
    assert relerr(xg, g['x_global']) <= 1e-4
    for r in (4, 8, 16):
        assert relerr(feats[r], g[f'feat{r}']) <= 1e-4, r
    for r, v in feats.items():
        st = g[f'feat{r}_stats']
        assert abs(v.std() - st[1]) <= 1e-4 * st[1] and abs(np.abs(v).max() - st[2]) <= 1e-4 * st[2], r
    assert err <= tol, err
    d = np.abs(comp.astype(int) - g['composite_u8'].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 2e-3


def test_generator_simt_gen128(golden):
    _check_generator('gen128_c64', 1, golden, 1e-3)


def test_generator_tc_gen128(golden):
    _check_generator('gen128_c64', 0, golden, 1e-3)


def test_generator_tc_halo_gen128(golden):
    _check_generator('gen128_c64', 3, golden, 1e-3)    # every layer (4x4 .. 128x128, down-2, up-2 passes) on the halo kernel


def test_generator_tc_tap_gen128(golden):
    _check_generator('gen128_c64', 2, golden, 1e-3)


def test_generator_tc_gen256(golden):
    _check_generator('gen256', 0, golden, 1e-3)        # north_star: within 1e-3 max-abs of the reference


def test_generator_tc_halo_gen256(golden):
    _check_generator('gen256', 3, golden, 1e-3)


def test_generator_tc_pair_gen256(golden):
    _check_generator('gen256', 4, golden, 1e-3)        # every Co % 256 == 0 layer (plain, stride-2, up-2 passes) on cta_group::2


def test_generator_tc_gen512(golden):
    _check_generator('gen512', 0, golden, 1e-3)


def test_generator_tc_vs_simt_batch_and_random_noise():
    """Full-size 256^2 model, batch 3: the tensor-core path must agree with the fp32 FMA kernel on identical
    operands, including noise_mode='random' under a fixed seed and batch-size independence of each sample."""
    sd = O.synthetic_state_dict(256, seed=3)
    G = H.build_generator(256, sd, device=DEV)
    x, z = O.synthetic_inputs(3, 256, seed=3)
    outs = {}
    for impl in (1, 0, 3):
        G.engine(impl=impl)
        torch.manual_seed(123)
        outs[impl] = G(t(x), t(z), None, noise_mode='random').cpu().numpy()
    assert np.abs(outs[0] - outs[1]).max() <= 1e-3, np.abs(outs[0] - outs[1]).max()
    assert np.abs(outs[3] - outs[1]).max() <= 1e-3, np.abs(outs[3] - outs[1]).max()
    G.engine(impl=0)
    torch.manual_seed(123)
    again = G(t(x), t(z), None, noise_mode='random').cpu().numpy()
    assert np.array_equal(again, outs[0])
    one = G(t(x[:1]), t(z[:1]), None, noise_mode='const').cpu().numpy()
    three = G(t(x), t(z), None, noise_mode='const').cpu().numpy()
    # the batch-global style normaliser (stylegan.py:147) cancels in the demodulation up to its 1e-8 epsilon
    assert np.abs(one[0] - three[0]).max() <= 1e-3 * max(1.0, np.abs(three).max())


def test_generator_tc_ragged_and_empty_batch():
    """The last batch of an eval run is ragged (shgan_default.py:269-274 feeds whatever the sampler returns): batch 1 against the
    oracle, and an empty batch returns an empty image like the reference's PyTorch ops do."""
    sd = O.synthetic_state_dict(128, seed=11, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd, 8192, 64, device=DEV)
    x, z = O.synthetic_inputs(3, 128, seed=11)
    for n in (1, 3):
        img = G(t(x[:n]), t(z[:n]), None, noise_mode='const').cpu().numpy()
        ref = O.generator(sd, x[:n], z[:n], 128)
        assert img.shape == ref.shape and np.abs(img - ref).max() <= 1e-3
    img0 = G(t(x[:0]), t(z[:0]), None, noise_mode='const')
    assert tuple(img0.shape) == (0, 3, 128, 128)
    img0, comp0 = G.forward_composite(t(x[:0]), t(z[:0]), noise_mode='random')
    assert tuple(img0.shape) == (0, 3, 128, 128) and tuple(comp0.shape) == (0, 3, 128, 128) and comp0.dtype == torch.uint8


def test_generator_tc_state_dict_reload_and_deepcopy():
    import copy
    sd1 = O.synthetic_state_dict(128, seed=1, ch_base=8192, ch_max=64)
    sd2 = O.synthetic_state_dict(128, seed=2, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd1, 8192, 64, device=DEV)
    x, z = O.synthetic_inputs(2, 128, seed=1)
    a = G(t(x), t(z), None, noise_mode='const').cpu().numpy()
    G2 = copy.deepcopy(G)
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd2.items()}, strict=True)   # must invalidate the packed operands
    b = G(t(x), t(z), None, noise_mode='const').cpu().numpy()
    assert np.abs(a - b).max() > 1e-2
    assert np.array_equal(G2(t(x), t(z), None, noise_mode='const').cpu().numpy(), a)
    ref = O.generator(sd2, x, z, 128)
    assert np.abs(b - ref).max() <= 1e-3


# ------------------------------------------------------------------------------- discriminator
@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('case', DISCRIMINATOR_CASES, ids=[c[0] for c in DISCRIMINATOR_CASES])
def test_discriminator_golden(case, impl, golden):
    name, res, chb, chm, batch, seed = case
    D = H.build_discriminator(res, O.synthetic_discriminator_state_dict(res, seed=seed, ch_base=chb, ch_max=chm), chb, chm, device=DEV)
    D.engine(impl=impl)
    x, _ = O.synthetic_inputs(batch, res, seed=seed)
    y = D(t(x), None).cpu().numpy()
    ref = golden(name)['out']
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), np.abs(y - ref).max()


def test_generator_then_discriminator_tc_step():
    """C4-style step: G forward -> composite -> D(cat[mask - 0.5, fake]) on the tensor-core path vs the oracle."""
    sd_g = O.synthetic_state_dict(128, seed=5, ch_base=8192, ch_max=64)
    sd_d = O.synthetic_discriminator_state_dict(128, seed=5, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd_g, 8192, 64, device=DEV)
    D = H.build_discriminator(128, sd_d, 8192, 64, device=DEV)
    x, z = O.synthetic_inputs(4, 128, seed=5)
    img = G(t(x), t(z), None, noise_mode='const')
    m = t(x)[:, 0:1] + 0.5
    fake = t(x)[:, 1:4] * m + img * (1 - m)
    logits = D(torch.cat([t(x)[:, 0:1], fake], dim=1), None).cpu().numpy()
    img_ref = O.generator(sd_g, x, z, 128)
    mr = x[:, 0:1] + 0.5
    fake_ref = x[:, 1:4] * mr + img_ref * (1 - mr)
    ref = O.discriminator(sd_d, np.concatenate([x[:, 0:1], fake_ref], axis=1).astype(np.float32), 128)
    assert np.abs(logits - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------------------- fused up-sampling convolution
def _up2_reference(x, w, fy, fx, gain):
    """fp64 torch: conv_transpose2d(stride 2) with the un-flipped weights, then the separable blur with zero pad 1."""
    n, c, h, wd = x.shape
    z = torch.nn.functional.conv_transpose2d(x.double(), w.double().permute(1, 0, 2, 3), stride=2)      # [n,co,2h+1,2w+1]
    f = torch.outer(torch.tensor(fy, dtype=torch.float64), torch.tensor(fx, dtype=torch.float64)).to(x.device) * gain
    co = z.shape[1]
    return torch.nn.functional.conv2d(z, f[None, None].expand(co, 1, 4, 4), padding=1, groups=co)        # correlation, as applied


UP2_SHAPES = [(2, 64, 64, 8, 8), (1, 128, 64, 20, 23), (2, 64, 128, 13, 7), (1, 256, 192, 4, 4), (3, 64, 64, 1, 2), (2, 192, 64, 33, 40),
              (1, 512, 128, 16, 16), (1, 384, 64, 9, 9)]


@pytest.mark.parametrize('shape', UP2_SHAPES, ids=[str(s) for s in UP2_SHAPES])
@pytest.mark.parametrize('passes', [3, 1])
@pytest.mark.parametrize('narrow', [False, True], ids=['auto', 'narrow'])
@pytest.mark.parametrize('cluster', [False, True], ids=['single', 'pair'])
def test_conv_up2_tc_vs_fp64(shape, passes, narrow, cluster):
    """Both epilogue instances of the fused kernel (C <= 256 runs the 16-warp one unless `narrow` forces the 8-warp one), each as
    single CTAs and as two-CTA clusters sharing every weight load by TMA multicast (odd tile counts leave one CTA of the last
    pair with an invalid tile)."""
    from shgan_b200 import kernels as K, packing as P
    n, ci, co, h, wd = shape
    if narrow and ci > 256:
        pytest.skip('C > 256 always runs the 8-warp epilogue')
    g = torch.Generator().manual_seed(ci * 7 + co + h)
    x = torch.randn(n, ci, h, wd, generator=g).to(DEV)
    w = (torch.randn(co, ci, 3, 3, generator=g) / (3 * ci ** 0.5)).to(DEV)
    fy, fx = [0.25, 0.75, 0.75, 0.25], [0.125, 0.375, 0.375, 0.125]
    fy = [1.0, 2.5, 3.0, 0.5]                                  # asymmetric taps: catches any flip / transposition of the blur
    fx = [0.5, 3.0, 2.0, 1.5]
    ref = _up2_reference(x, w, fy, fx, 4.0 / 49.0)
    xp = K.nchw_to_planes(x)
    uh, ul = P.pack_up2_weight(w)
    # raw accumulator path (identity epilogue)
    y32 = torch.empty((n, 2 * h, 2 * wd, co), device=DEV)
    K.conv_up2(xp, uh, ul, fx, fy, 4.0 / 49.0, K.make_epilogue(out_f32=y32), passes=passes, narrow=narrow, cluster=cluster)
    tol = 1e-5 if passes == 3 else 5e-3
    assert relerr(y32.permute(0, 3, 1, 2).cpu().numpy(), ref.cpu().numpy()) <= tol
    # full epilogue: demod, per-sample noise, bias, lrelu + clamp, skip, next-layer style; planes + fp32 outputs
    dc = (0.5 + torch.rand(n, co, generator=g)).to(DEV)
    nzv = torch.randn(n, 1, 2 * h, 2 * wd, generator=g).to(DEV)
    strength = torch.tensor(0.3, device=DEV)
    bias = torch.randn(co, generator=g).to(DEV)
    skip = torch.randn(n, co, 2 * h, 2 * wd, generator=g).to(DEV)
    ns = (0.5 + torch.rand(n, co, generator=g)).to(DEV)
    out = K.Planes.empty(n, 2 * h, 2 * wd, co, DEV)
    K.conv_up2(xp, uh, ul, fx, fy, 4.0 / 49.0,
               K.make_epilogue(dcoef=dc, wgain=0.7, noise=nzv, noise_sn=4 * h * wd, noise_strength=strength, bias=bias, act=True,
                               act_alpha=0.2, act_gain=2 ** 0.5, act_clamp=3.0, skip=K.nchw_to_planes(skip), next_scale=ns, out=out,
                               out_f32=y32), passes=passes, narrow=narrow, cluster=cluster)
    v = ref * dc[:, :, None, None] * 0.7 + nzv * 0.3 + bias[None, :, None, None]
    v = (torch.where(v >= 0, v, v * 0.2) * 2 ** 0.5).clamp(-3.0, 3.0)
    sk = K.planes_to_nchw(K.nchw_to_planes(skip)).double()
    v = v + sk
    assert relerr(y32.permute(0, 3, 1, 2).cpu().numpy(), v.cpu().numpy()) <= (2e-5 if passes == 3 else 5e-3)
    got = K.planes_to_nchw(out).double()
    assert relerr(got.cpu().numpy(), (v * ns[:, :, None, None]).cpu().numpy()) <= (2e-5 if passes == 3 else 5e-3)


def test_generator_tc_fused_up2_equals_unfused():
    """The fused up-sampling launch against the 4 RAW passes + blur kernel it replaces, on the full 256^2 model."""
    sd = O.synthetic_state_dict(256, seed=4)
    G = H.build_generator(256, sd, device=DEV)
    x, z = O.synthetic_inputs(2, 256, seed=4)
    eng = G.engine(impl=0)
    a = G(t(x), t(z), None, noise_mode='const').cpu().numpy()
    eng.fuse_up2 = False
    eng._graphs = {}
    b = G(t(x), t(z), None, noise_mode='const').cpu().numpy()
    eng.fuse_up2 = True
    assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_generator_tc_forward_inpaint_equals_prepared_input():
    """Input preparation fused into fromrgb (shgan_default.py:269-274): forward_inpaint(real, mask, z) must reproduce
    forward_composite(cat([mask - 0.5, real * mask]), z) bit for bit, eagerly and under graph replay."""
    sd = O.synthetic_state_dict(128, seed=11, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd, 8192, 64, device=DEV)
    x, z = O.synthetic_inputs(3, 128, seed=11)
    mask = t(x[:, 0:1] + 0.5)
    real = torch.rand(3, 3, 128, 128, generator=torch.Generator().manual_seed(3)).to(DEV) * 2 - 1
    xin = torch.cat([mask - 0.5, real * mask], dim=1)
    for _ in range(2):
        img_a, comp_a = G.forward_composite(xin, t(z), noise_mode='const')
        img_b, comp_b = G.forward_inpaint(real, mask, t(z), noise_mode='const')
        assert torch.equal(img_a, img_b) and torch.equal(comp_a, comp_b)
