"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through ref_import.py) on seeded inputs.  Build-container only:
the GPU box never runs this, it only reads the committed *.npz files.

    python tests/golden/make_golden.py [--only NAME]

Each fixture stores the reference OUTPUTS (and small inputs); large inputs / weights are
regenerated on the fly from `oracle.shgan_oracle.synthetic_state_dict / synthetic_inputs`,
which are pure numpy-PCG64 and therefore identical on every machine.
The script also prints the oracle-vs-reference error for every fixture as it writes it.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from oracle import shgan_oracle as O  # noqa: E402

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def save(name, **arrs):
    path = os.path.join(HERE, name + '.npz')
    np.savez(path, **arrs)
    print(f'  wrote {name}.npz  ({os.path.getsize(path) / 1e6:.2f} MB)')


def err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


UPFIRDN_CASES = [
    # name, shape, filter, kwargs          (the first four are exactly the variants the model uses)
    ('blur_pad2', (2, 3, 16, 16), 'f1331', dict(padding=[2, 2, 2, 2])),                      # conv2d_resample.py:117-120 (down path pre-blur)
    ('blur_pad1_gain4', (2, 3, 17, 17), 'f1331', dict(padding=[1, 1, 1, 1], gain=4)),         # conv2d_resample.py:139 (up path post-blur)
    ('up2_rgb', (2, 3, 8, 8), 'f1331', dict(up=2, padding=[2, 1, 2, 1], gain=4)),             # upsample2d, comodgan.py:331-332
    ('down2_pad1', (2, 3, 16, 16), 'f1331', dict(down=2, padding=[1, 1, 1, 1])),              # downsample2d / discriminator skip
    ('asym_filter_noflip', (1, 2, 9, 11), 'asym', dict(padding=[2, 1, 0, 3])),
    ('asym_filter_flip', (1, 2, 9, 11), 'asym', dict(padding=[2, 1, 0, 3], flip_filter=True)),
    ('up2_down2_asym', (1, 2, 7, 6), 'asym', dict(up=2, down=2, padding=[3, 2, 1, 2], gain=2.5)),
    ('up3x_down2y', (1, 1, 5, 6), 'asym', dict(up=[3, 1], down=[1, 2], padding=[1, 2, 2, 1])),
    ('negative_pad_crop', (1, 2, 12, 12), 'f1331', dict(padding=[-1, 2, 3, -2])),
    ('identity_filter_none', (1, 2, 6, 6), None, dict(up=2, padding=[0, 1, 0, 1])),
    ('sep_1d_8tap', (1, 2, 20, 20), 'sep8', dict(padding=[4, 3, 4, 3])),
    ('single_pixel', (1, 1, 1, 1), 'f1331', dict(padding=[2, 1, 2, 1], up=2, gain=4)),
]


def get_filter(kind):
    if kind is None:
        return None
    if kind == 'f1331':
        return O.setup_filter([1, 3, 3, 1])
    if kind == 'asym':
        return np.array([[1, 2, 3, 4, 5], [0, -1, 2, 0.5, 1], [3, 1, -2, 1, 0.25]], np.float32) / 7
    if kind == 'sep8':
        return O.setup_filter([1, 2, 3, 4, 4, 3, 2, 1])
    raise KeyError(kind)


def gen_upfirdn2d(R):
    out = {}
    for i, (name, shape, fk, kw) in enumerate(UPFIRDN_CASES):
        x = rng(100 + i).standard_normal(shape).astype(np.float32)
        f = get_filter(fk)
        y = R.upfirdn2d.upfirdn2d(t(x), None if f is None else t(f), impl='ref', **kw).numpy()
        yo = O.upfirdn2d(x, f, **kw)
        print(f'  upfirdn2d/{name}: oracle err {err(y, yo):.2e}  out {y.shape}')
        out[name] = y
    # setup_filter goldens
    for nm, args, kw in [('sf_1331', [1, 3, 3, 1], {}), ('sf_8tap', [1, 2, 3, 4, 4, 3, 2, 1], {}),
                         ('sf_gain_flip', [1, 2, 3], dict(flip_filter=True, gain=4)), ('sf_2d', [[1, 2], [3, 4]], dict(normalize=False))]:
        f = R.upfirdn2d.setup_filter(args, **kw).numpy()
        print(f'  setup_filter/{nm}: oracle err {err(f, O.setup_filter(args, **kw)):.2e}')
        out[nm] = f
    save('upfirdn2d', **out)


CONV_CASES = [
    ('plain3x3', dict(ci=8, co=12, k=3, hw=10, up=1, down=1, flip_weight=True)),
    ('down2_3x3', dict(ci=8, co=12, k=3, hw=12, up=1, down=2, flip_weight=True)),
    ('up2_3x3', dict(ci=8, co=12, k=3, hw=7, up=2, down=1, flip_weight=False)),
    ('plain1x1', dict(ci=8, co=3, k=1, hw=9, up=1, down=1, flip_weight=True)),
    ('down2_1x1', dict(ci=8, co=6, k=1, hw=12, up=1, down=2, flip_weight=True)),
]


def gen_conv2d_resample(R):
    out = {}
    f = O.setup_filter([1, 3, 3, 1])
    for i, (name, c) in enumerate(CONV_CASES):
        g = rng(200 + i)
        x = g.standard_normal((2, c['ci'], c['hw'], c['hw'])).astype(np.float32)
        w = g.standard_normal((c['co'], c['ci'], c['k'], c['k'])).astype(np.float32)
        ff = f if (c['up'] > 1 or c['down'] > 1) else None
        y = R.conv2d_resample.conv2d_resample(t(x), t(w), f=None if ff is None else t(ff), up=c['up'], down=c['down'],
                                              padding=c['k'] // 2, flip_weight=c['flip_weight']).numpy()
        yo = O.conv2d_resample(x, w, f=ff, up=c['up'], down=c['down'], padding=c['k'] // 2, flip_weight=c['flip_weight'])
        print(f'  conv2d_resample/{name}: oracle err {err(y, yo):.2e} out {y.shape}')
        out[name] = y
    save('conv2d_resample', **out)


MODCONV_CASES = [
    ('demod_3x3', dict(ci=16, co=24, k=3, hw=8, up=1, demod=True, noise=True)),
    ('demod_up2', dict(ci=16, co=24, k=3, hw=8, up=2, demod=True, noise=True)),
    ('torgb_1x1', dict(ci=16, co=3, k=1, hw=8, up=1, demod=False, noise=False)),
    ('demod_3x3_c64', dict(ci=64, co=64, k=3, hw=16, up=1, demod=True, noise=True)),
    ('demod_up2_c64', dict(ci=64, co=64, k=3, hw=16, up=2, demod=True, noise=True)),
]


def modconv_inputs(i, c, batch=3):
    g = rng(300 + i)
    x = g.standard_normal((batch, c['ci'], c['hw'], c['hw'])).astype(np.float32)
    w = g.standard_normal((c['co'], c['ci'], c['k'], c['k'])).astype(np.float32)
    s = (1 + 0.5 * g.standard_normal((batch, c['ci']))).astype(np.float32)
    res = c['hw'] * c['up']
    nz = (0.1 * g.standard_normal((batch, 1, res, res))).astype(np.float32) if c['noise'] else None
    return x, w, s, nz


def gen_modulated_conv2d(R):
    out = {}
    f = O.setup_filter([1, 3, 3, 1])
    for i, (name, c) in enumerate(MODCONV_CASES):
        x, w, s, nz = modconv_inputs(i, c)
        ff = f if c['up'] > 1 else None
        kw = dict(up=c['up'], padding=c['k'] // 2, demodulate=c['demod'], flip_weight=(c['up'] == 1))
        y = R.stylegan.modulated_conv2d(t(x), t(w), t(s), noise=None if nz is None else t(nz),
                                        resample_filter=None if ff is None else t(ff), fused_modconv=True, **kw).numpy()
        y2 = R.stylegan.modulated_conv2d(t(x), t(w), t(s), noise=None if nz is None else t(nz),
                                         resample_filter=None if ff is None else t(ff), fused_modconv=False, **kw).numpy()
        yo = O.modulated_conv2d(x, w, s, noise=nz, resample_filter=ff, **kw)
        print(f'  modulated_conv2d/{name}: oracle err {err(y, yo):.2e}  ref fused-vs-unfused {err(y, y2):.2e}  |y|max {np.abs(y).max():.2f}')
        out[name] = y
    save('modulated_conv2d', **out)


def gen_small_ops(R):
    out = {}
    g = rng(400)
    x = (g.standard_normal((4, 7, 5, 5)) * 100).astype(np.float32)
    act = R.utils.get_unit()(ACT)()
    out['lrelu_x'] = x
    out['lrelu_y'] = act(t(x.copy())).numpy()
    out['lrelu_y_gain'] = act(t(x.copy()), gain=float(np.sqrt(0.5))).numpy()
    print('  lrelu_agc: oracle err', err(out['lrelu_y'], O.lrelu_agc(x)), err(out['lrelu_y_gain'], O.lrelu_agc(x, extra_gain=float(np.sqrt(0.5)))))
    # dense / small mapping network
    torch.manual_seed(0)
    m = R.stylegan.Mapping(z_dim=64, c_dim=0, w_dim=64, num_ws=5, num_layers=8, activation=ACT, lr_multiplier=0.01).eval()
    sd = {'mapping.' + k: v.numpy() for k, v in m.state_dict().items()}
    z = g.standard_normal((3, 64)).astype(np.float32)
    with torch.no_grad():
        ws = m(t(z), None).numpy()
    for k, v in sd.items():
        out['map_' + k] = v
    out['map_z'] = z
    out['map_ws'] = ws
    print('  mapping: oracle err', err(ws, O.mapping(sd, z, 5)))
    d = R.stylegan.dense(40, 24, bias=True, bias_init=1, activation=ACT, lr_multi=0.5)
    xin = g.standard_normal((5, 40)).astype(np.float32)
    with torch.no_grad():
        yd = d(t(xin)).numpy()
    out['dense_w'] = d.weight.detach().numpy()
    out['dense_b'] = d.bias.detach().numpy()
    out['dense_x'] = xin
    out['dense_y'] = yd
    print('  dense: oracle err', err(yd, O.dense(xin, out['dense_w'], out['dense_b'], lr_multi=0.5, act=True)))
    save('small_ops', **out)


def gen_shu(R):
    out = {}
    sd = {k: v for k, v in O.synthetic_state_dict(256, seed=7).items() if k.startswith('encoder.shu')}
    shu = R.shgan.SHU(32, 32, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=64, lowest_res=4,
                      tail_sigma_mult=3, gaussian_at_input_res=False).eval()
    shu.load_state_dict({k[len('encoder.shu.'):]: t(v) for k, v in sd.items()}, strict=True)
    x = rng(500).standard_normal((1, 32, 64, 64)).astype(np.float32)
    stages = {}
    h1 = shu.conv0.register_forward_hook(lambda m, i, o: stages.__setitem__('conv0', o.detach().numpy().copy()))
    h2 = shu.df1.register_forward_hook(lambda m, i, o: stages.__setitem__('df1', o.detach().numpy().copy()))
    with torch.no_grad():
        y = shu(t(x))
    h1.remove(); h2.remove()
    yo, so = O.shu_forward(sd, x, return_stages=True)
    for r in y:
        print(f'  shu/out{r}: oracle err {err(y[r].numpy(), yo[r]):.2e} |y|max {y[r].abs().max():.3f}')
        out[f'out{r}'] = y[r].numpy()
    print(f"  shu/conv0 err {err(stages['conv0'], so['conv0']):.2e}  df1 err {err(stages['df1'], so['df1']):.2e}")
    out['stage_conv0'] = stages['conv0'].astype(np.float16)  # stage tensors kept in half to bound fixture size
    out['stage_df1'] = stages['df1']
    cw = R.shgan.make_cweight([2, 3], [64, 33]).numpy()
    out['cweight_2x3_64x33'] = cw
    print('  make_cweight 64x33: oracle err', err(cw, O.make_cweight([2, 3], [64, 33])))
    for hs, ws_ in [(16, 9), (8, 5), (7, 5)]:
        cwx = R.shgan.make_cweight([2, 3], [hs, ws_]).numpy()
        out[f'cweight_2x3_{hs}x{ws_}'] = cwx
        print(f'  make_cweight {hs}x{ws_}: oracle err', err(cwx, O.make_cweight([2, 3], [hs, ws_])))
    gm = O.gaussian_weight_maps()
    for r, v in shu.gaussian_weight_map.items():
        out[f'gauss{r}'] = v.numpy()
        print(f'  gaussian map {r}: oracle err {err(v.numpy(), gm[r]):.2e}')
    # C5-style sweep sizes: SHU at other input resolutions (outputs only for the two lowest bands to stay small)
    for n in (16, 128):
        shu_n = R.shgan.SHU(32, 32, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=n, lowest_res=4).eval()
        shu_n.load_state_dict({k[len('encoder.shu.'):]: t(v) for k, v in sd.items()}, strict=True)
        xn = rng(510 + n).standard_normal((1, 32, n, n)).astype(np.float32)
        with torch.no_grad():
            yn = shu_n(t(xn))
        yon = O.shu_forward(sd, xn, input_res=n)
        print(f'  shu@{n}: oracle err', max(err(yn[r].numpy(), yon[r]) for r in yn))
        out[f'sweep{n}_out4'] = yn[4].numpy()
        out[f'sweep{n}_out{n}_sum'] = np.array([float(yn[n].double().sum()), float(yn[n].double().abs().sum())])
    save('shu', **out)


build_reference_generator = ref_import.build_reference_generator


GENERATOR_CASES = [
    # name, resolution, ch_base, ch_max, batch, seed
    ('gen128_c64', 128, 8192, 64, 2, 11),
    ('gen256', 256, 32768, 512, 1, 12),
    ('gen512', 512, 32768, 512, 1, 13),
]


def gen_generator(R, only=None):
    for name, res, chb, chm, batch, seed in GENERATOR_CASES:
        if only and only != name:
            continue
        sd = O.synthetic_state_dict(res, seed=seed, ch_base=chb, ch_max=chm)
        G = build_reference_generator(R, res, chb, chm)
        G.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)  # also pins the key/shape contract
        x, z = O.synthetic_inputs(batch, res, seed=seed)
        inter = {}
        hooks = [G.mapping.register_forward_hook(lambda m, i, o: inter.__setitem__('ws', o.numpy().copy())),
                 G.encoder.register_forward_hook(lambda m, i, o: inter.__setitem__('enc', o))]
        with torch.no_grad():
            img = G(t(x), t(z), torch.zeros(batch, 0), noise_mode='const').numpy()
        for h in hooks:
            h.remove()
        x_global, feats = inter['enc']
        out = dict(img=img, ws0=inter['ws'][:, 0], x_global=x_global.numpy())
        for r, v in feats.items():
            v = v.numpy()
            out[f'feat{r}_stats'] = np.array([v.mean(), v.std(), np.abs(v).max()], np.float64)
            if r <= 16:
                out[f'feat{r}'] = v
        u8 = (t(x)[:, 1:4] * (t(x)[:, 0:1] + 0.5) + t(img) * (1 - (t(x)[:, 0:1] + 0.5)))
        u8 = (u8 * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).numpy()
        out['composite_u8'] = u8
        print(f'  {name}: |img|max {np.abs(img).max():.3f}; running oracle ...', flush=True)
        io, into = O.generator(sd, x, z, res, return_intermediates=True)
        print(f'  {name}: oracle err img {err(img, io):.2e}  x_global {err(out["x_global"], into["x_global"]):.2e} '
              f'ws {err(out["ws0"], into["ws"][:, 0]):.2e}  composite mismatches {(O.composite_uint8(x, io) != u8).sum()}')
        save(name, **out)


DISCRIMINATOR_CASES = [
    # name, resolution, ch_base, ch_max, batch, seed
    ('disc128_c64', 128, 8192, 64, 4, 21),
    ('disc256', 256, 32768, 512, 8, 22),
]


def gen_discriminator(R):
    for name, res, chb, chm, batch, seed in DISCRIMINATOR_CASES:
        sd = O.synthetic_discriminator_state_dict(res, seed=seed, ch_base=chb, ch_max=chm)
        D = R.comodgan.Discriminator(resolution=res, ic_n=4, ch_base=chb, ch_max=chm, use_fp16_before_res=None).eval().requires_grad_(False)
        D.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        assert list(D.state_dict().keys()) == [k for k, _ in O.discriminator_state_dict_spec(res, ch_base=chb, ch_max=chm)]
        x, _ = O.synthetic_inputs(batch, res, seed=seed)
        inter = {}
        hooks = []
        for r in (8, 16):
            blk = getattr(D, f'b{r}')
            hooks.append(blk.register_forward_hook(lambda m, i, o, r=r: inter.__setitem__(r // 2, o[0].numpy().copy())))
        with torch.no_grad():
            y = D(t(x), None).numpy()
        for h in hooks:
            h.remove()
        yo, io = O.discriminator(sd, x, res, return_intermediates=True)
        print(f'  {name}: out {y.ravel()[:4]}  oracle err {err(y, yo):.2e}  x4 err {err(inter[4], io[4]):.2e}')
        save(name, out=y, x4=inter[4], x8=inter[8])


# The configurations bench.py actually times (BASELINE.json configs C2 / C3): the batch changes the tile schedule
# (multi-image tiles, the odd last tile pair of the two-SM kernel) and the batch-global style normaliser
# (stylegan.py:147), so they get their own fixtures.  A [B,3,R,R] fp32 image set is 50 MB at C3: the fixture keeps a
# strided sub-sample whose phase differs per sample (sample n keeps pixels (n%4 + 4i, (n//4)%4 + 4j), so the 16 phases of
# the 4x4 lattice are all covered) plus per-sample statistics of the full images.
BENCH_CASES = [
    # name, resolution, batch, seed
    ('gen512_b16', 512, 16, 14),
    ('gen256_b32', 256, 32, 15),
]
BENCH_STRIDE = 4


def bench_subsample(img):
    st = BENCH_STRIDE
    return np.stack([img[n, :, (n % st)::st, ((n // st) % st)::st] for n in range(img.shape[0])])


def gen_bench_configs(R, only=None):
    for name, res, batch, seed in BENCH_CASES:
        if only and only != name:
            continue
        sd = O.synthetic_state_dict(res, seed=seed)
        G = build_reference_generator(R, res)
        G.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        x, z = O.synthetic_inputs(batch, res, seed=seed)
        torch.set_num_threads(os.cpu_count() or 1)
        imgs = []
        with torch.no_grad():
            # ONE call over the whole batch: the style normaliser is batch-global
            img = G(t(x), t(z), torch.zeros(batch, 0), noise_mode='const').numpy()
        xt, it = t(x), t(img)
        u8 = (xt[:, 1:4] * (xt[:, 0:1] + 0.5) + it * (1 - (xt[:, 0:1] + 0.5)))
        u8 = (u8 * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).numpy()
        stats = np.stack([[v.mean(), v.std(), np.abs(v).max(), np.abs(v).astype(np.float64).sum()] for v in img]).astype(np.float64)
        print(f'  {name}: |img|max {np.abs(img).max():.3f}')
        save(name, img_sub=bench_subsample(img), composite_sub=bench_subsample(u8), stats=stats)


def gen_discriminator512(R):
    name, res, batch, seed = 'disc512', 512, 8, 23
    sd = O.synthetic_discriminator_state_dict(res, seed=seed)
    D = ref_import.build_reference_discriminator(R, res)
    D.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    x, _ = O.synthetic_inputs(batch, res, seed=seed)
    with torch.no_grad():
        y = D(t(x), None).numpy()
    print(f'  {name}: out {y.ravel()[:4]}')
    save(name, out=y)


MASK_CASES = [(0, 512, (0.0, 1.0)), (1, 256, (0.0, 1.0)), (2, 512, (0.2, 0.6))]      # np.random seed, size, hole_range


def gen_random_mask():
    """RandomMask / RandomBrush (lib/data_factory/ds_ffhq.py:145-217) run from the reference source under
    np.random.seed(seed): 3 consecutive masks per case, bit-packed.  (The two functions are exec'd out of the file because
    importing lib.data_factory pulls in packages this image does not have.)"""
    import math
    from PIL import Image, ImageDraw
    src = open(os.path.join(ref_import.reference_root(), 'lib', 'data_factory', 'ds_ffhq.py')).read()
    ns = dict(math=math, np=np, Image=Image, ImageDraw=ImageDraw)
    exec(src[src.index('def RandomBrush('):src.index('###############\n# ffhq_simple')], ns)
    out = {}
    for seed, size, hr in MASK_CASES:
        np.random.seed(seed)
        for k in range(3):
            m = ns['RandomMask'](size, list(hr))
            out[f'seed{seed}_s{size}_{k}'] = np.packbits(m[0].astype(np.uint8))
            print(f'  mask seed {seed} size {size} #{k}: hole ratio {1 - m.mean():.3f}')
    save('random_mask', **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    R = ref_import.import_reference()
    steps = dict(upfirdn2d=gen_upfirdn2d, conv2d_resample=gen_conv2d_resample, modulated_conv2d=gen_modulated_conv2d,
                 small_ops=gen_small_ops, shu=gen_shu)
    for nm, fn in steps.items():
        if args.only in (None, nm):
            print(nm)
            fn(R)
    if args.only in (None, 'discriminator'):
        print('discriminator')
        gen_discriminator(R)
    if args.only in (None, 'random_mask'):
        print('random_mask')
        gen_random_mask()
    if args.only in (None, 'disc512'):
        print('discriminator 512')
        gen_discriminator512(R)
    if args.only is None or (args.only.startswith('gen') and not args.only.endswith(('_b16', '_b32'))):
        print('generator')
        gen_generator(R, None if args.only in (None, 'generator') else args.only)
    if args.only in (None, 'bench') or (args.only or '').endswith(('_b16', '_b32')):
        print('benchmark configurations')
        gen_bench_configs(R, None if args.only in (None, 'bench') else args.only)


if __name__ == '__main__':
    main()
