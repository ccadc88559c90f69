"""GPU parity tests (run with `-m gpu` on the B200 box): every kernel is called through the C ABI (ctypes) and
compared with the CPU oracle / the golden fixtures generated from the reference (tests/golden/make_golden.py).

Tolerances: the reference path is fp32; all kernels here accumulate in fp32 (convolutions: fp16 hi/lo split
operands, fp32 TMEM accumulation, dropped term < 2^-22), so op-level results must agree with the reference to
fp32 rounding: 2e-5 relative to the output scale unless stated otherwise.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import shgan_oracle as O  # noqa: E402  (checker only)
from golden.make_golden import UPFIRDN_CASES, CONV_CASES, MODCONV_CASES, get_filter, modconv_inputs, rng  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


# ---------------------------------------------------------------------------------------------- upfirdn2d
@pytest.mark.parametrize('case', UPFIRDN_CASES, ids=[c[0] for c in UPFIRDN_CASES])
def test_upfirdn2d_golden(case, golden):
    from shgan_b200 import ops
    name, shape, fk, kw = case
    i = [c[0] for c in UPFIRDN_CASES].index(name)
    x = rng(100 + i).standard_normal(shape).astype(np.float32)
    f = get_filter(fk)
    y = ops.upfirdn2d(t(x), None if f is None else t(f), **kw).cpu().numpy()
    ref = golden('upfirdn2d')[name]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_upfirdn2d_wrappers_and_sizes():
    from shgan_b200 import ops
    f = O.setup_filter([1, 3, 3, 1])
    g = np.random.default_rng(0)
    for shape in [(2, 5, 33, 47), (1, 3, 64, 64), (3, 2, 129, 130)]:
        x = g.standard_normal(shape).astype(np.float32)
        for fn, ofn, kw in [(ops.upsample2d, O.upsample2d, {}), (ops.downsample2d, O.downsample2d, {}),
                            (ops.filter2d, O.filter2d, dict(gain=2.0)),
                            (ops.upfirdn2d, O.upfirdn2d, dict(padding=[2, 2, 2, 2])),
                            (ops.upfirdn2d, O.upfirdn2d, dict(padding=[1, 1, 1, 1], gain=4))]:
            y = fn(t(x), t(f), **kw).cpu().numpy()
            ref = ofn(x, f, **kw)
            assert y.shape == ref.shape
            assert np.abs(y - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (fn.__name__, shape)


@pytest.mark.parametrize('kind', ['rank1_asym', 'full_rank'])
@pytest.mark.parametrize('flip', [False, True])
def test_upfirdn2d_4x4_stride1_kernels(kind, flip):
    """up == down == 1 with a 4x4 filter: outer products run the row-walking separable kernel, anything else the tile kernel
    (the choice is made on the device); several column blocks / row segments, ragged widths (scalar-store path), crops."""
    from shgan_b200 import ops
    g = np.random.default_rng(7)
    if kind == 'rank1_asym':
        f = np.outer([1.0, 2.0, 3.0, 4.0], [4.0, -1.0, 2.0, 0.5]).astype(np.float32) / 20
    else:
        f = g.standard_normal((4, 4)).astype(np.float32)
    for shape, pad in [((2, 3, 70, 130), [2, 2, 2, 2]), ((1, 2, 150, 257), [1, 1, 1, 1]), ((1, 1, 66, 300), [-1, 3, 0, -2]),
                       ((3, 1, 5, 9), [3, 3, 3, 3]), ((1, 2, 131, 512), [2, 1, 1, 2])]:
        x = g.standard_normal(shape).astype(np.float32)
        y = ops.upfirdn2d(t(x), t(f), padding=pad, flip_filter=flip, gain=1.5).cpu().numpy()
        ref = O.upfirdn2d(x, f, padding=pad, flip_filter=flip, gain=1.5)
        assert y.shape == ref.shape
        assert np.abs(y - ref).max() <= 3e-6 * max(1.0, np.abs(ref).max()), (kind, flip, shape, pad)


def test_upfirdn2d_errors():
    from shgan_b200 import ops
    f = t(O.setup_filter([1, 3, 3, 1]))
    with pytest.raises(RuntimeError):
        ops.upfirdn2d(torch.zeros(1, 1, 4, 4), f.cpu())            # no CPU path
    with pytest.raises(RuntimeError):
        ops.upfirdn2d(torch.zeros(1, 1, 2, 2, device=DEV), f, padding=[-3, -3, -3, -3])   # output < 1x1
    y = ops.upfirdn2d(torch.zeros(0, 3, 8, 8, device=DEV), f, padding=1)  # empty batch
    assert y.shape == (0, 3, 7, 7)


# ---------------------------------------------------------------------------------------------- layout
def test_layout_roundtrip():
    from shgan_b200 import kernels as K
    g = np.random.default_rng(1)
    x = (g.standard_normal((3, 40, 9, 13)) * 10).astype(np.float32)
    p = K.nchw_to_planes(t(x))
    back = K.planes_to_nchw(p).cpu().numpy()
    assert np.abs(back - x).max() <= 1e-6 * np.abs(x).max()
    # slice add: planes[:, :, :, 8:24] += y
    y = g.standard_normal((3, 16, 9, 13)).astype(np.float32)
    K.planes_add_nchw(p, t(y), 8)
    ref = x.copy()
    ref[:, 8:24] += y
    back = K.planes_to_nchw(p).cpu().numpy()
    assert np.abs(back - ref).max() <= 2e-6 * np.abs(ref).max()
    sl = K.planes_to_nchw(p, c_off=8, c=16).cpu().numpy()
    assert np.abs(sl - ref[:, 8:24]).max() <= 2e-6 * np.abs(ref).max()
    # add + scale
    sc = g.standard_normal((3, 40)).astype(np.float32)
    q = K.nchw_to_planes(t(x), add=p, scale=t(sc))
    ref2 = (x + ref) * sc[:, :, None, None]
    assert np.abs(K.planes_to_nchw(q).cpu().numpy() - ref2).max() <= 3e-6 * np.abs(ref2).max()
    z = g.standard_normal((2, 7, 5, 24)).astype(np.float32)
    assert np.array_equal(K.nhwc_to_nchw_f32(t(z)).cpu().numpy(), z.transpose(0, 3, 1, 2))


# ---------------------------------------------------------------------------------------------- small ops
def test_dense_mapping_golden(golden):
    from shgan_b200 import kernels as K
    g = golden('small_ops')
    out = torch.empty((5, 24), device=DEV)
    K.dense(t(g['dense_x']), t(g['dense_w']), t(g['dense_b']), out, 0.5 / np.sqrt(40), 0.5, True, 0.2, np.sqrt(2), 256.0)
    assert relerr(out.cpu().numpy(), g['dense_y']) <= 2e-6
    x = K.normalize_2nd_moment(t(g['map_z']))
    for i in range(8):
        w, b = t(g[f'map_mapping.fc{i}.weight']), t(g[f'map_mapping.fc{i}.bias'])
        o = torch.empty((3, 64), device=DEV)
        x = K.dense(x, w, b, o, 0.01 / np.sqrt(64), 0.01, True, 0.2, np.sqrt(2), 256.0)
    assert relerr(x.cpu().numpy(), g['map_ws'][:, 0]) <= 1e-5
    # concatenated / strided inputs (styles = affine(cat[w, x_global]))
    r = np.random.default_rng(3)
    a, b2, w = r.standard_normal((9, 4, 32)).astype(np.float32), r.standard_normal((9, 64)).astype(np.float32), r.standard_normal((20, 96)).astype(np.float32)
    o = torch.empty((9, 20), device=DEV)
    K.dense(t(a)[:, 2], t(w), None, o, 0.3, 1.0, False, x1=t(b2))
    ref = np.concatenate([a[:, 2], b2], 1) @ w.T * 0.3
    assert relerr(o.cpu().numpy(), ref) <= 2e-6


@pytest.mark.parametrize('shape', [(16, 8192, 1024, 8192), (19, 1024, 100, 1024), (3, 512, 512, 512), (16, 1536, 64, 512), (8, 2048, 40, 1024),
                                   (33, 768, 17, 768)], ids=str)
def test_dense_cluster_split(shape):
    """Layers whose output count does not fill the GPU split the input features over a thread-block cluster (partial sums reduced
    through distributed shared memory in a fixed order): against an fp64 matmul, twice (bit-identical: deterministic)."""
    from shgan_b200 import kernels as K
    b, i, o, i0 = shape
    g = torch.Generator().manual_seed(b * 31 + o)
    x0 = torch.randn(b, i0, generator=g).to(DEV)
    x1 = torch.randn(b, i - i0, generator=g).to(DEV) if i > i0 else None
    w = torch.randn(o, i, generator=g).to(DEV)
    bias = torch.randn(o, generator=g).to(DEV)
    y = torch.empty((b, o), device=DEV)
    K.dense(x0, w, bias, y, 1.0 / np.sqrt(i), 0.7, True, 0.2, np.sqrt(2), 256.0, x1=x1)
    xx = torch.cat([x0, x1], 1) if x1 is not None else x0
    ref = xx.double() @ w.double().T / np.sqrt(i) + bias.double() * 0.7
    ref = torch.where(ref >= 0, ref, ref * 0.2) * np.sqrt(2)
    assert relerr(y.cpu().numpy(), ref.cpu().numpy()) <= 3e-6
    y2 = torch.empty_like(y)
    K.dense(x0, w, bias, y2, 1.0 / np.sqrt(i), 0.7, True, 0.2, np.sqrt(2), 256.0, x1=x1)
    assert torch.equal(y, y2)


def test_style_prep():
    from shgan_b200 import kernels as K
    r = np.random.default_rng(4)
    s = (1 + 0.5 * r.standard_normal((5, 128))).astype(np.float32)
    wsq = r.random((64, 128)).astype(np.float32)
    s_hat, dc = torch.empty((5, 128), device=DEV), torch.empty((5, 64), device=DEV)
    K.style_prep(t(s), t(wsq), s_hat, dc, True)
    sn = s / np.sqrt(np.mean(s.astype(np.float64) ** 2))
    assert relerr(s_hat.cpu().numpy(), sn) <= 2e-6
    assert relerr(dc.cpu().numpy(), 1 / np.sqrt((sn ** 2) @ wsq.T + 1e-8)) <= 3e-6
    K.style_prep(t(s), None, s_hat, None, False, 0.25)
    assert relerr(s_hat.cpu().numpy(), s * 0.25) <= 1e-7


def test_style_prep_batched():
    """All style layers in one launch (columns of one concatenated dense output) == the per-layer entry point."""
    from shgan_b200 import kernels as K
    r = np.random.default_rng(41)
    specs = [(512, 512, True, 1.0), (512, 3, False, 0.044), (64, 128, True, 1.0), (256, 3, False, 0.0625), (128, 64, True, 1.0)]
    n, total = 7, sum(c for c, _, _, _ in specs)
    raw = (1 + 0.5 * r.standard_normal((n, total + 8))).astype(np.float32)      # row stride > total
    raw_t = t(raw)[:, :total]
    layers, refs, off = [], [], 0
    for ci, co, demod, ps in specs:
        wsq = t(r.random((co, ci)).astype(np.float32)) if demod else None
        L = dict(offset=off, ci=ci, co=co, demod=demod, pre_scale=ps, wsq=wsq, s_hat=torch.empty((n, ci), device=DEV),
                 dcoef=torch.empty((n, co), device=DEV) if demod else None)
        sh, dc = torch.empty((n, ci), device=DEV), (torch.empty((n, co), device=DEV) if demod else None)
        K.style_prep(raw_t[:, off:off + ci].contiguous(), wsq, sh, dc, demod, ps)
        layers.append(L); refs.append((sh, dc)); off += ci
    K.style_prep_batched(raw_t, layers)
    for L, (sh, dc) in zip(layers, refs):
        assert relerr(L['s_hat'].cpu().numpy(), sh.cpu().numpy()) <= 1e-6
        if dc is not None:
            assert relerr(L['dcoef'].cpu().numpy(), dc.cpu().numpy()) <= 2e-6


# ---------------------------------------------------------------------------------------------- FIR / pointwise
def test_fir_nhwc_epilogue_and_parity():
    from shgan_b200 import kernels as K
    r = np.random.default_rng(5)
    f = O.setup_filter([1, 3, 3, 1])
    x = r.standard_normal((2, 16, 11, 13)).astype(np.float32)          # NCHW
    skip = r.standard_normal((2, 16, 12, 14)).astype(np.float32)
    bias = r.standard_normal(16).astype(np.float32)
    dco = r.random((2, 16)).astype(np.float32) + 0.5
    nxt = r.random((2, 16)).astype(np.float32) + 0.5
    nz = r.standard_normal((12, 14)).astype(np.float32)
    xp = K.nchw_to_planes(t(x))
    out = K.Planes.empty(2, 12, 14, 16, DEV)
    of32 = torch.empty((2, 12, 14, 16), device=DEV)
    strength = torch.tensor(0.3, device=DEV)
    # the epilogue struct holds raw device pointers: keep every operand alive until the launch has been issued
    ops_ = dict(dcoef=t(dco), noise=t(nz), bias=t(bias), skip=K.nchw_to_planes(t(skip)), next_scale=t(nxt), f=t(f))
    epi = K.make_epilogue(dcoef=ops_['dcoef'], noise=ops_['noise'], noise_sn=0, noise_strength=strength, bias=ops_['bias'], act=True,
                          act_alpha=0.2, act_gain=np.sqrt(2), act_clamp=256.0, skip=ops_['skip'], next_scale=ops_['next_scale'],
                          out=out, out_f32=of32)
    K.fir_nhwc(xp, ops_['f'], 4.0, (2, 2, 2, 2), epi)
    ref = O.upfirdn2d(x, f, padding=[2, 2, 2, 2], gain=4) * dco[:, :, None, None] + nz[None, None] * 0.3 + bias[None, :, None, None]
    ref = O.lrelu_agc(ref) + skip
    assert relerr(K.nhwc_to_nchw_f32(of32).cpu().numpy(), ref) <= 3e-6
    assert relerr(K.planes_to_nchw(out).cpu().numpy(), ref * nxt[:, :, None, None]) <= 3e-6
    # fp32 NHWC input + parity-split output
    xin = torch.from_numpy(x.transpose(0, 2, 3, 1).copy()).to(DEV)
    ph, pw = 6, 7
    par = K.Planes.empty(4 * 2, ph, pw, 16, DEV)
    K.fir_nhwc(xin, t(f), 1.0, (2, 2, 2, 2), K.make_epilogue(out=par), parity_split=True)
    ref = O.upfirdn2d(x, f, padding=[2, 2, 2, 2])
    got = K.planes_to_nchw(par).cpu().numpy().reshape(4, 2, 16, ph, pw)
    for py in range(2):
        for px in range(2):
            sub = ref[:, :, py::2, px::2]
            assert relerr(got[py * 2 + px][:, :, :sub.shape[2], :sub.shape[3]], sub) <= 3e-6


@pytest.mark.parametrize('parity', [0, 1, 2])
@pytest.mark.parametrize('shape,pads', [((2, 32, 11, 13), (2, 2, 2, 2)), ((1, 64, 70, 37), (2, 2, 2, 2)), ((3, 96, 40, 66), (1, 1, 1, 1)),
                                        ((1, 128, 33, 9), (2, 1, 0, 3)), ((2, 64, 4, 4), (2, 2, 2, 2))], ids=str)
def test_fir_walk_kernel(shape, pads, parity):
    """The row-walking blur (planes -> planes, identity epilogue, rank-1 hint: the launch in front of every stride-2 conv) against
    the oracle and, bit for bit up to fp32 summation order, against the two-phase kernel it replaces; asymmetric separable taps."""
    from shgan_b200 import kernels as K
    r = np.random.default_rng(shape[1] + shape[2])
    f = np.outer([1.0, 2.0, 3.5, 0.5], [0.5, 3.0, 2.0, 1.5]).astype(np.float32) / 49
    x = r.standard_normal(shape).astype(np.float32)
    n, c, h, w = shape
    oh, ow = h + pads[2] + pads[3] - 3, w + pads[0] + pads[1] - 3
    xp = K.nchw_to_planes(t(x))
    ft = t(f)
    ref = O.upfirdn2d(x, f, padding=list(pads), flip_filter=True, gain=1.25)       # fir_nhwc applies f as given (correlation)
    outs = []
    for two_phase in (False, True):
        if parity == 0:
            out = K.Planes.empty(n, oh, ow, c, DEV)
        elif parity == 1:
            out = K.Planes.empty(4 * n, (oh + 1) // 2, (ow + 1) // 2, c, DEV)
        else:
            out = K.Planes.empty(n, (oh + 1) // 2, (ow + 1) // 2, c, DEV)
        out.hi.zero_(); out.lo.zero_()
        K.fir_nhwc(xp, ft, 1.25, pads, K.make_epilogue(out=out), parity_split=parity, rank1=True, two_phase=two_phase)
        outs.append(K.planes_to_nchw(out).cpu().numpy())
    got = outs[0]
    if parity == 0:
        assert relerr(got, ref) <= 3e-6
    elif parity == 2:
        assert relerr(got, ref[:, :, ::2, ::2]) <= 3e-6
    else:
        g4 = got.reshape(4, n, c, (oh + 1) // 2, (ow + 1) // 2)
        for py in range(2):
            for px in range(2):
                sub = ref[:, :, py::2, px::2]
                assert relerr(g4[py * 2 + px][:, :, :sub.shape[2], :sub.shape[3]], sub) <= 3e-6
    assert relerr(outs[0], outs[1]) <= 3e-6          # same cells written (the rest stays zero in both), same values


def test_fromrgb_and_torgb_combine():
    from shgan_b200 import kernels as K
    r = np.random.default_rng(6)
    x = r.standard_normal((2, 4, 10, 12)).astype(np.float32)
    w = r.standard_normal((64, 4)).astype(np.float32)
    b = r.standard_normal(64).astype(np.float32)
    out = K.fromrgb(t(x), t(w), t(b), 0.5, 0.2, np.sqrt(2), 256.0, K.Planes.empty(2, 10, 12, 64, DEV))
    ref = O.lrelu_agc(np.einsum('nchw,oc->nohw', x, w * 0.5) + b[None, :, None, None])
    assert relerr(K.planes_to_nchw(out).cpu().numpy(), ref) <= 3e-6
    f = O.setup_filter([1, 3, 3, 1])
    prev = r.standard_normal((2, 3, 5, 6)).astype(np.float32)
    part = r.standard_normal((2, 10, 12, 3, 4)).astype(np.float32)
    bias = r.standard_normal(3).astype(np.float32)
    cx = np.concatenate([(r.random((2, 1, 10, 12)) > 0.5).astype(np.float32) - 0.5, r.uniform(-1, 1, (2, 3, 10, 12)).astype(np.float32)], 1)
    img = torch.empty((2, 3, 10, 12), device=DEV)
    comp = torch.empty((2, 3, 10, 12), dtype=torch.uint8, device=DEV)
    K.torgb_combine(t(prev), t(part), t(bias), t(f), img, comp_x=t(cx), comp_out=comp)
    ref = O.upsample2d(prev, f) + part[..., :3].sum(3).transpose(0, 3, 1, 2) + bias[None, :, None, None]
    assert relerr(img.cpu().numpy(), ref) <= 3e-6
    refc = O.composite_uint8(cx, img.cpu().numpy())
    assert np.abs(comp.cpu().numpy().astype(int) - refc.astype(int)).max() <= 1
    assert (comp.cpu().numpy() != refc).mean() < 1e-3
    img0 = torch.empty((2, 3, 10, 12), device=DEV)
    K.torgb_combine(None, t(part), t(bias), None, img0)
    assert relerr(img0.cpu().numpy(), part[..., :3].sum(3).transpose(0, 3, 1, 2) + bias[None, :, None, None]) <= 2e-6


# ---------------------------------------------------------------------------------------------- SHU
def test_shu_golden(golden):
    from shgan_b200.model_zoo.shgan import SHU
    g = golden('shu')
    sd = {k: v for k, v in O.synthetic_state_dict(256, seed=7).items() if k.startswith('encoder.shu')}
    shu = SHU(32, 32, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=64, lowest_res=4)
    shu.load_state_dict({k[len('encoder.shu.'):]: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    shu = shu.to(DEV)
    x = rng(500).standard_normal((1, 32, 64, 64)).astype(np.float32)
    y = shu(t(x))
    assert sorted(y) == [4, 8, 16, 32, 64]
    for r in y:
        ref = g[f'out{r}']
        assert y[r].shape == ref.shape
        assert relerr(y[r].cpu().numpy(), ref) <= 2e-5, r
    # other input resolutions (C5 sweep sizes) and batch > 1 against the oracle
    for n in (16, 128):
        shu_n = SHU(32, 32, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=n, lowest_res=4)
        shu_n.load_state_dict(shu.state_dict())
        shu_n = shu_n.to(DEV)
        xn = rng(510 + n).standard_normal((1, 32, n, n)).astype(np.float32)
        yn = shu_n(t(xn))
        assert relerr(yn[4].cpu().numpy(), g[f'sweep{n}_out4']) <= 2e-5
        tot = g[f'sweep{n}_out{n}_sum']
        assert abs(float(yn[n].double().abs().sum()) - tot[1]) <= 1e-4 * tot[1]
    xb = rng(7).standard_normal((3, 32, 64, 64)).astype(np.float32)
    yb = shu(t(xb))
    ref = O.shu_forward(sd, xb)
    for r in ref:
        assert relerr(yb[r].cpu().numpy(), ref[r]) <= 2e-5


@pytest.mark.parametrize('n_in', [256, 512, 4, 8, 32])
def test_shu_large_and_small_input_res_vs_oracle(n_in):
    """BASELINE.json config C5 sweeps the unit over input_res 4..512: sizes above 128 run their transforms as row / column
    passes through global memory, sizes up to 32 (and every band of at most 32 x 32) as one register-resident transform per thread
    (shu_small.cu); checked against the oracle (numpy pocketfft)."""
    from shgan_b200.model_zoo.shgan import SHU
    sd = {k: v for k, v in O.synthetic_state_dict(256, seed=7).items() if k.startswith('encoder.shu')}
    shu = SHU(32, 32, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=n_in, lowest_res=4)
    shu.load_state_dict({k[len('encoder.shu.'):]: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    shu = shu.to(DEV)
    x = rng(520 + n_in).standard_normal((2, 32, n_in, n_in)).astype(np.float32)
    y = shu(t(x))
    ref = O.shu_forward(sd, x, input_res=n_in)
    assert sorted(y) == sorted(ref)
    for r in ref:
        assert relerr(y[r].cpu().numpy(), ref[r]) <= 2e-5, r


def test_shu_narrow_uses_fma_mix():
    """C != 32: the fp32-FMA channel-mix kernel (the tensor-core mix needs 2C == 64) against the oracle."""
    from shgan_b200.model_zoo.shgan import SHU
    g = rng(77)
    c = 16
    sd = {'s.conv0.weight': (g.standard_normal((2 * c, 2 * c, 1, 1)) / 6).astype(np.float32),
          's.conv0.bias': (0.1 * g.standard_normal(2 * c)).astype(np.float32),
          's.df1.weight': (1 / 64 + 0.1 / 64 * g.standard_normal((2 * c, 2 * c * 6))).astype(np.float32)}
    shu = SHU(c, c, dfilter_freedom=[2, 3], dfilter_type='piecewise_linear', input_res=32, lowest_res=4)
    shu.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    x = g.standard_normal((2, c, 32, 32)).astype(np.float32)
    y = shu.to(DEV)(t(x))
    ref = O.shu_forward(sd, x, prefix='s', input_res=32, lowest_res=4)
    for r in ref:
        assert relerr(y[r].cpu().numpy(), ref[r]) <= 2e-5, r


def _shu_numpy(x, conv0_w, conv0_b, df1_w, cw, masks, r):
    """float64 restatement of SHU.forward (shgan.py:312-336) for arbitrary cw / Gaussian masks (numpy pocketfft)."""
    n = x.shape[0]
    f = np.fft.rfft2(x.astype(np.float64), norm='forward')
    f = np.concatenate([f[:, :, r // 2 + 1:], f[:, :, :r // 2 + 1]], axis=2)
    s1 = np.concatenate([f.real, f.imag], axis=1)
    tt = np.maximum(np.einsum('oi,nisk->nosk', conv0_w.astype(np.float64), s1) + conv0_b.astype(np.float64)[None, :, None, None], 0)
    y = np.einsum('io,nisk->nosk', df1_w.astype(np.float64), tt).reshape(n, 64, 6, r, r // 2 + 1)
    s2 = (y * cw.astype(np.float64)[None, None]).sum(2)
    fc = s2[:, :32] + 1j * s2[:, 32:]
    out = {}
    for k, mk in masks.items():
        sp = fc[:, :, r // 2 - k // 2:r // 2 + k // 2, :k // 2 + 1] * mk.astype(np.float64)[None, None]
        sp = np.concatenate([sp[:, :, k - k // 2 - 1:], sp[:, :, :k - k // 2 - 1]], axis=2)
        out[k] = np.fft.irfft2(sp, s=(k, k), norm='forward')
    return out


@pytest.mark.parametrize('lowest,dense_cw,prepacked', [(4, False, True), (8, False, False), (16, True, True), (64, True, False)])
def test_shu_r64_register_fft_path(lowest, dense_cw, prepacked):
    """input_res 64 (the released model): register-resident radix-8 transforms + tcgen05 channel mix (shu_fft64.cu,
    shu_mix_tc.cu) through the C ABI, against float64 numpy: every lowest_res (bands that are not produced), a cw with all six
    anchors active everywhere (three anchor pairs per tile instead of two), random Gaussian masks, weights packed per call."""
    from shgan_b200 import kernels as K, packing as P
    g = rng(900 + lowest)
    r, n = 64, 5
    conv0_w = (g.standard_normal((64, 64)) / 8).astype(np.float32)
    conv0_b = (0.1 * g.standard_normal(64)).astype(np.float32)
    df1_w = (1 / 64 + 0.1 / 64 * g.standard_normal((64, 384))).astype(np.float32)
    cw = g.uniform(0.1, 1.0, (6, r, r // 2 + 1)).astype(np.float32) if dense_cw else P.make_cweight((2, 3), (r, r // 2 + 1)).numpy()
    reslist = [k for k in (4, 8, 16, 32, 64) if k >= lowest]
    masks = {k: g.uniform(0.2, 1.0, (k, k // 2 + 1)).astype(np.float32) for k in reslist}
    x = g.standard_normal((n, 32, r, r)).astype(np.float32)
    gauss = np.concatenate([masks[k].reshape(-1) for k in reslist])
    outs = [torch.empty(n, 32, k, k, device=DEV) for k in reslist]
    packed = K.shu_pack(t(conv0_w), t(df1_w)) if prepacked else None
    K.shu_fwd(t(x), t(conv0_w), t(conv0_b), t(df1_w), t(cw), t(gauss), outs, lowest, packed=packed)
    ref = _shu_numpy(x, conv0_w, conv0_b, df1_w, cw, masks, r)
    for o, k in zip(outs, reslist):
        assert relerr(o.cpu().numpy(), ref[k]) <= 2e-5, k


def test_shu_r64_unaligned_input_falls_back():
    """The bulk copies of the register-FFT path need 16-byte aligned planes; an x that is only 4-byte aligned takes the generic
    shared-memory transforms (row-major spectrum) with the same results."""
    from shgan_b200 import kernels as K, packing as P
    g = rng(77)
    r, n = 64, 2
    conv0_w = (g.standard_normal((64, 64)) / 8).astype(np.float32)
    conv0_b = (0.1 * g.standard_normal(64)).astype(np.float32)
    df1_w = (1 / 64 + 0.1 / 64 * g.standard_normal((64, 384))).astype(np.float32)
    cw = P.make_cweight((2, 3), (r, r // 2 + 1)).numpy()
    mk = {k: v.numpy() for k, v in P.gaussian_band_masks(r, 4, 3, False).items()}
    reslist = sorted(mk)
    x = g.standard_normal((n, 32, r, r)).astype(np.float32)
    buf = torch.empty(x.size + 1, device=DEV)
    xd = buf[1:].view(n, 32, r, r)
    xd.copy_(t(x))
    assert xd.data_ptr() % 16 == 4
    gauss = np.concatenate([mk[k].reshape(-1) for k in reslist])
    outs = [torch.empty(n, 32, k, k, device=DEV) for k in reslist]
    K.shu_fwd(xd, t(conv0_w), t(conv0_b), t(df1_w), t(cw), t(gauss), outs, 4)
    ref = _shu_numpy(x, conv0_w, conv0_b, df1_w, cw, mk, r)
    for o, k in zip(outs, reslist):
        assert relerr(o.cpu().numpy(), ref[k]) <= 2e-5, k


def test_fir_decimate_and_mbstd():
    from shgan_b200 import kernels as K
    r = np.random.default_rng(8)
    f = O.setup_filter([1, 3, 3, 1])
    x = r.standard_normal((3, 32, 12, 10)).astype(np.float32)
    xp = K.nchw_to_planes(t(x))
    out = K.Planes.empty(3, 6, 5, 32, DEV)
    ft = t(f)
    K.fir_nhwc(xp, ft, 1.0, (1, 1, 1, 1), K.make_epilogue(out=out), parity_split=2)
    ref = O.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1])
    assert relerr(K.planes_to_nchw(out).cpu().numpy(), ref) <= 3e-6
    # minibatch std (group 4, 1 channel) fused with the concat and channel padding
    xb = r.standard_normal((8, 64, 4, 4)).astype(np.float32)
    pb = K.nchw_to_planes(t(xb))
    ob = K.mbstd_append(pb, K.Planes.empty(8, 4, 4, 128, DEV), 4)
    got = K.planes_to_nchw(ob).cpu().numpy()
    refb = O.minibatch_std(xb, 4, 1)
    assert relerr(got[:, :65], refb) <= 3e-6 and np.abs(got[:, 65:]).max() == 0


def test_prepare_input_and_composite_cat():
    """The two eval-loop glue kernels against the reference expressions (shgan_default.py:269-274 and :257-260)."""
    from shgan_b200 import kernels as K
    g = torch.Generator().manual_seed(9)
    real = (torch.rand(3, 3, 64, 48, generator=g) * 2 - 1).to(DEV)
    mask = (torch.rand(3, 1, 64, 48, generator=g) > 0.4).float().to(DEV)
    x = K.prepare_input(real, mask)
    assert torch.equal(x, torch.cat([mask - 0.5, real * mask], dim=1))
    img = torch.randn(3, 3, 64, 48, generator=g).to(DEV) * 3
    m = x[:, 0:1] + 0.5
    ref = torch.cat([x[:, 0:1], x[:, 1:4] * m + img * (1 - m)], dim=1)
    assert torch.equal(K.composite_cat(x, img), ref)
