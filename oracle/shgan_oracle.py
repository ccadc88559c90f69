"""CPU oracle for the SH-GAN generator-forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (SHI-Labs/SH-GAN @ a9ba83c5,
`lib/model_zoo/**`).  It is NOT part of the product: only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and there only
as the checker / the CPU baseline.  The product path (`shgan_b200`) never imports this module
and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (imports /root/reference unmodified) and committed under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against them.

Every function cites the reference file:line it restates.  All arrays are NCHW numpy arrays;
`dtype` may be float32 (bit-for-bit comparable with the reference up to summation order) or
float64 (the arbiter used to judge the CUDA kernels' split-precision error).

The conv primitive has two backends: 'numpy' (tap-wise BLAS matmuls; default, used by all
tests) and 'torch' (torch.nn.functional.conv2d on CPU threads -- the call the reference itself
makes on CPU via conv2d_gradfix.py:38,43 -- used only to time the CPU baseline in bench.py).
"""
import math

import numpy as np

_CONV_BACKEND = 'numpy'


def set_conv_backend(name):
    global _CONV_BACKEND
    assert name in ('numpy', 'torch')
    _CONV_BACKEND = name


# ----------------------------------------------------------------------------------------
# primitives: conv2d / conv_transpose2d  (torch.nn.functional semantics, groups=1)
# reference call sites: stylegan_utils/conv2d_gradfix.py:35-43
# ----------------------------------------------------------------------------------------

def conv2d(x, w, stride=1, padding=0):
    """Cross-correlation, NCHW x [N,Ci,H,W], w [Co,Ci,kh,kw] -> [N,Co,Ho,Wo]."""
    if _CONV_BACKEND == 'torch':
        import torch
        y = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x)),
                                       torch.from_numpy(np.ascontiguousarray(w)),
                                       stride=stride, padding=padding)
        return y.numpy()
    n, ci, h, wd = x.shape
    co, ci2, kh, kw = w.shape
    assert ci == ci2
    if padding:
        x = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    hp, wp = x.shape[2], x.shape[3]
    ho = (hp - kh) // stride + 1
    wo = (wp - kw) // stride + 1
    y = np.zeros((n, co, ho * wo), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            win = x[:, :, ky:ky + (ho - 1) * stride + 1:stride, kx:kx + (wo - 1) * stride + 1:stride]
            win = np.ascontiguousarray(win).reshape(n, ci, ho * wo)
            y += np.matmul(w[None, :, :, ky, kx], win)
    return y.reshape(n, co, ho, wo)


def conv_transpose2d(x, w, stride=1, padding=0):
    """torch.nn.functional.conv_transpose2d semantics, w [Ci,Co,kh,kw] -> [N,Co,(H-1)s+kh-2p, ...]."""
    if _CONV_BACKEND == 'torch':
        import torch
        y = torch.nn.functional.conv_transpose2d(torch.from_numpy(np.ascontiguousarray(x)),
                                                 torch.from_numpy(np.ascontiguousarray(w)),
                                                 stride=stride, padding=padding)
        return y.numpy()
    n, ci, h, wd = x.shape
    ci2, co, kh, kw = w.shape
    assert ci == ci2
    ho = (h - 1) * stride + kh
    wo = (wd - 1) * stride + kw
    y = np.zeros((n, co, ho, wo), dtype=x.dtype)
    xf = x.reshape(n, ci, h * wd)
    for ky in range(kh):
        for kx in range(kw):
            contrib = np.matmul(w[:, :, ky, kx].T[None], xf).reshape(n, co, h, wd)
            y[:, :, ky:ky + (h - 1) * stride + 1:stride, kx:kx + (wd - 1) * stride + 1:stride] += contrib
    if padding:
        y = y[:, :, padding:ho - padding, padding:wo - padding]
    return y


# ----------------------------------------------------------------------------------------
# upfirdn2d  (stylegan_utils/upfirdn2d.py)
# ----------------------------------------------------------------------------------------

def _parse_scaling(s):  # upfirdn2d.py:33-40
    if isinstance(s, int):
        s = [s, s]
    sx, sy = s
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(p):  # upfirdn2d.py:42-51
    if isinstance(p, int):
        p = [p, p]
    p = list(p)
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    return p


def _get_filter_size(f):  # upfirdn2d.py:53-64
    if f is None:
        return 1, 1
    return f.shape[-1], f.shape[0]


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:66-92.  [1,3,3,1] has <8 taps, so it is expanded to the full 4x4 outer product."""
    if f is None:
        f = 1
    f = np.asarray(f, dtype=np.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.size >= 8)
    if f.ndim == 1 and not separable:
        f = np.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f[tuple(slice(None, None, -1) for _ in range(f.ndim))]
    f = f * (gain ** (f.ndim / 2))
    return np.ascontiguousarray(f.astype(np.float32))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:98-138 (`_upfirdn2d_ref`, also what the CUDA plugin upfirdn2d.cu:97-200 computes):
    zero-insert upsample, pad/crop, correlate with the flipped filter, decimate, scale by gain."""
    n, c, h, w = x.shape
    if f is None:
        f = np.ones([1, 1], dtype=np.float32)
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    z = np.zeros((n, c, h, upy, w, upx), dtype=x.dtype)
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(n, c, h * upy, w * upx)
    z = np.pad(z, ((0, 0), (0, 0), (max(pady0, 0), max(pady1, 0)), (max(padx0, 0), max(padx1, 0))))
    z = z[:, :, max(-pady0, 0):z.shape[2] - max(-pady1, 0), max(-padx0, 0):z.shape[3] - max(-padx1, 0)]
    f = (f * (gain ** (f.ndim / 2))).astype(x.dtype)
    if not flip_filter:
        f = f[tuple(slice(None, None, -1) for _ in range(f.ndim))]
    if f.ndim == 1:
        fs = [f[None, :], f[:, None]]
    else:
        fs = [f]
    for ff in fs:
        fh, fw = ff.shape
        ho, wo = z.shape[2] - fh + 1, z.shape[3] - fw + 1
        acc = np.zeros((n, c, ho, wo), dtype=x.dtype)
        for i in range(fh):
            for j in range(fw):
                acc += ff[i, j] * z[:, :, i:i + ho, j:j + wo]
        z = acc
    return np.ascontiguousarray(z[:, :, ::downy, ::downx])


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):  # upfirdn2d.py:279-314
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2,
         pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):  # upfirdn2d.py:316-351
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2,
         pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1):  # upfirdn2d.py:245-277
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


# ----------------------------------------------------------------------------------------
# conv2d_resample  (stylegan_utils/conv2d_resample.py:57-154), groups == 1
# ----------------------------------------------------------------------------------------

def _conv2d_wrapper(x, w, stride=1, padding=0, transpose=False, flip_weight=True):
    # conv2d_resample.py:26-51
    if not flip_weight:
        w = w[:, :, ::-1, ::-1]
    if isinstance(padding, (list, tuple)):
        assert padding[0] == padding[1]
        padding = padding[0]
    if transpose:
        return conv_transpose2d(x, w, stride=stride, padding=padding)
    return conv2d(x, w, stride=stride, padding=padding)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, flip_weight=True, flip_filter=False):
    co, ci, kh, kw = w.shape
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:  # :93-97
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:  # :98-102
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and (down > 1 and up == 1):  # :105-108
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and (up > 1 and down == 1):  # :111-114
        x = _conv2d_wrapper(x, w, flip_weight=flip_weight)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:  # :117-120
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, stride=down, flip_weight=flip_weight)
    if up > 1:  # :123-142
        w = np.transpose(w, (1, 0, 2, 3))
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv2d_wrapper(x, w, stride=up, padding=[pyt, pxt], transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1:  # :145-147
        if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
            assert px0 == py0
            return _conv2d_wrapper(x, w, padding=px0, flip_weight=flip_weight)
    raise NotImplementedError('generic fallback conv2d_resample.py:150-154 is not reached by this model')


# ----------------------------------------------------------------------------------------
# activations / dense  (common/utils.py:117-146, stylegan.py:66-101)
# ----------------------------------------------------------------------------------------

SQRT2 = float(np.sqrt(2))


def lrelu_agc(x, alpha=0.2, gain=SQRT2, clamp=256, extra_gain=1):
    """common/utils.py:135-143: leaky_relu(alpha) -> * (gain*extra_gain) -> clamp(+-clamp*extra_gain)."""
    x = np.where(x >= 0, x, x * x.dtype.type(alpha))
    act_gain = gain * extra_gain
    if act_gain != 1:
        x = x * x.dtype.type(act_gain)
    if clamp is not None:
        c = x.dtype.type(clamp * extra_gain)
        x = np.clip(x, -c, c)
    return x


def dense(x, weight, bias, lr_multi=1, act=False):
    """stylegan.py:87-98.  weight [out,in]; weight_gain = lr_multi/sqrt(in); bias_gain = lr_multi."""
    dt = x.dtype
    wg = dt.type(lr_multi / np.sqrt(weight.shape[1]))
    w = weight.astype(dt) * wg
    y = x @ w.T
    if bias is not None:
        b = bias.astype(dt)
        if lr_multi != 1:
            b = b * dt.type(lr_multi)
        y = y + b[None]
    if act:
        y = lrelu_agc(y)
    return y


# ----------------------------------------------------------------------------------------
# modulated_conv2d  (stylegan.py:103-193)
# ----------------------------------------------------------------------------------------

def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    dt = x.dtype
    n = x.shape[0]
    co, ci, kh, kw = weight.shape
    weight = weight.astype(dt)
    styles = styles.astype(dt)
    w = None
    dcoefs = None
    if demodulate:  # :145-147  (weight pre-norm per out-channel, BATCH-GLOBAL style normaliser)
        weight = weight * (1 / np.sqrt(np.mean(np.square(weight), axis=(1, 2, 3), keepdims=True)))
        styles = styles * (1 / np.sqrt(np.mean(np.square(styles))))
    if demodulate or fused_modconv:  # :149-151
        w = weight[None] * styles.reshape(n, 1, -1, 1, 1)
    if demodulate:  # :155
        dcoefs = 1 / np.sqrt(np.sum(np.square(w), axis=(2, 3, 4)) + dt.type(1e-8))
    if demodulate and fused_modconv:  # :168-169
        w = w * dcoefs.reshape(n, -1, 1, 1, 1)
    if not fused_modconv:  # :172-181
        x = x * styles.reshape(n, -1, 1, 1)
        x = conv2d_resample(x, weight, f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate:
            x = x * dcoefs.reshape(n, -1, 1, 1)
        if noise is not None:
            x = x + noise.astype(dt)
        return x
    # fused: grouped conv with groups = batch  == per-sample conv (:184-192)
    outs = []
    for i in range(n):
        outs.append(conv2d_resample(x[i:i + 1], np.ascontiguousarray(w[i]), f=resample_filter, up=up, down=down,
                                    padding=padding, flip_weight=flip_weight))
    x = np.concatenate(outs, axis=0)
    if noise is not None:
        x = x + noise.astype(dt)
    return x


# ----------------------------------------------------------------------------------------
# layers operating on a flat state_dict (reference key names, Appendix A of SURVEY.md)
# ----------------------------------------------------------------------------------------

_F1331 = None


def _resample_filter():
    global _F1331
    if _F1331 is None:
        _F1331 = setup_filter([1, 3, 3, 1])
    return _F1331


def conv2d_layer(sd, prefix, x, up=1, down=1, act=True, gain=1):
    """stylegan.py:226-238."""
    dt = x.dtype
    w = sd[prefix + '.weight'].astype(dt)
    k = w.shape[2]
    w = w * dt.type(1 / np.sqrt(w.shape[1] * k * k))
    f = _resample_filter() if (up > 1 or down > 1) else None
    x = conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
    b = sd.get(prefix + '.bias')
    if b is not None:
        x = x + b.astype(dt).reshape(1, -1, 1, 1)
    if act:
        x = lrelu_agc(x, extra_gain=gain)
    else:
        x = x * dt.type(gain)
    return x


def synthesis_layer(sd, prefix, x, w_long, up=1, noise_mode='const', noise=None, fused_modconv=True):
    """stylegan.py:276-304.  noise_mode 'const' uses the noise_const buffer; 'random' takes `noise`
    ([N,1,res,res] standard normal, already drawn by the caller) ; 'none' adds nothing."""
    dt = x.dtype
    styles = dense(w_long, sd[prefix + '.affine.weight'], sd[prefix + '.affine.bias'])
    nz = None
    strength = sd[prefix + '.noise_strength'].astype(dt)
    if noise_mode == 'random':
        nz = noise.astype(dt) * strength
    elif noise_mode == 'const':
        nz = sd[prefix + '.noise_const'].astype(dt) * strength
    weight = sd[prefix + '.weight']
    f = _resample_filter() if up > 1 else None
    x = modulated_conv2d(x, weight, styles, noise=nz, up=up, padding=weight.shape[2] // 2, resample_filter=f,
                         flip_weight=(up == 1), fused_modconv=fused_modconv)
    x = x + sd[prefix + '.bias'].astype(dt).reshape(1, -1, 1, 1)
    return lrelu_agc(x)


def torgb_layer(sd, prefix, x, w_long, fused_modconv=True):
    """stylegan.py:325-337: styles * 1/sqrt(Cin), modconv 1x1 without demodulation, + bias."""
    dt = x.dtype
    weight = sd[prefix + '.weight']
    wg = dt.type(1 / np.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    styles = dense(w_long, sd[prefix + '.affine.weight'], sd[prefix + '.affine.bias']) * wg
    x = modulated_conv2d(x, weight, styles, demodulate=False, fused_modconv=fused_modconv)
    return x + sd[prefix + '.bias'].astype(dt).reshape(1, -1, 1, 1)


def mapping(sd, z, num_ws, num_layers=8, lr_multiplier=0.01, prefix='mapping'):
    """stylegan.py:394-430 with c_dim=0, truncation_psi=1 (eval path)."""
    dt = z.dtype
    x = z * (1 / np.sqrt(np.mean(np.square(z), axis=1, keepdims=True) + dt.type(1e-8)))  # :343-344
    for i in range(num_layers):
        x = dense(x, sd[f'{prefix}.fc{i}.weight'], sd[f'{prefix}.fc{i}.bias'], lr_multi=lr_multiplier, act=True)
    return np.repeat(x[:, None, :], num_ws, axis=1)


# ----------------------------------------------------------------------------------------
# Spectral Hint Unit  (shgan.py:70-160, 252-336)
# ----------------------------------------------------------------------------------------

def make_cweight(half_size, half_sample, dtype=np.float32):
    """Closed form of shgan.py:70-121 for type='piecewise_linear', oddeven_aligned=True.

    The reference bilinearly samples (grid_sample, align_corners=True, border padding) a one-hot
    lattice of h0 x w0 anchors, reflect-extended to the left.  Vertical sample positions map to
    lattice coordinate v_i = (i+1)/hs*(h0-1) for even hs (else i/(hs-1)*(h0-1)); horizontal ones to
    u_j = j/(ws-1)*(w0-1).  Bilinear sampling of a one-hot image is the product of two hat functions,
    so cw[r*w0+c, i, j] = hat(v_i - r) * hat(u_j - c)."""
    h0, w0 = half_size
    hs, ws = half_sample
    if hs % 2 == 0:
        hg = np.array([-1 + i / hs * 2 for i in range(hs + 1)])[1:]
    else:
        hg = np.array([-1 + i / (hs - 1) * 2 for i in range(hs)])
    wg = np.array([0 + i / (ws - 1) for i in range(ws)])
    hg = hg.astype(np.float32).astype(np.float64)  # the reference builds the grid as a float32 tensor
    wg = wg.astype(np.float32).astype(np.float64)
    v = (hg + 1) / 2 * (h0 - 1)
    u = wg * (w0 - 1)  # padded width 2*w0-1, x in [0,1] -> padded col (x+1)/2*(2*w0-2) = (w0-1) + x*(w0-1)

    def hat(t):
        return np.maximum(0.0, 1.0 - np.abs(t))
    cw = np.zeros((h0 * w0, hs, ws), dtype=np.float64)
    for r in range(h0):
        for c in range(w0):
            cw[r * w0 + c] = hat(v - r)[:, None] * hat(u - c)[None, :]
    return cw.astype(dtype)


def gaussian_weight_maps(input_res=64, lowest_res=4, tail_sigma_mult=3, gaussian_at_input_res=False):
    """shgan.py:280-310 (+ gaussian_heatmap_2d :162-250 with a single axis-aligned Gaussian; its
    3-sigma 'speedup' window covers the whole r x (r/2+1) grid for these sizes).  float64 maths,
    final cast to float32 exactly like `torch.Tensor(...).float()`."""
    reslist = [2 ** i for i in range(int(np.log2(lowest_res)), int(np.log2(input_res)) + 1)]
    rev = reslist[::-1]
    maps = {}

    def gauss(r):
        sigma = (r // 2) / tail_sigma_mult
        ci, cj = r // 2 - 1, 0
        hh = np.arange(r, dtype=np.float64)[:, None]
        ww = np.arange(r // 2 + 1, dtype=np.float64)[None, :]
        g = np.zeros((r, r // 2 + 1), dtype=np.float64)
        sr = int(3 * sigma + 1)
        h0_, h1_ = max(min(ci - sr, r), 0), max(min(ci + sr, r), 0)
        w0_, w1_ = max(min(cj - sr, r // 2 + 1), 0), max(min(cj + sr, r // 2 + 1), 0)
        inv = 1.0 / (sigma ** 2)
        e = np.exp(-0.5 * (((hh - ci) ** 2) * inv + ((ww - cj) ** 2) * inv))
        g[h0_:h1_, w0_:w1_] = np.maximum(g[h0_:h1_, w0_:w1_], e[h0_:h1_, w0_:w1_])
        return g
    for idx, r in enumerate(rev):
        if idx != 0:
            maps[r] = gauss(r)
            rp = rev[idx - 1]
            maps[rp][(rp // 2 - r // 2):(rp // 2 + r // 2), 0:(r // 2 + 1)] -= maps[r]
        elif gaussian_at_input_res:
            maps[r] = gauss(r)
        else:
            maps[r] = np.ones((r, r // 2 + 1), dtype=np.float32).astype(np.float64)
    return {r: maps[r].astype(np.float32) for r in reslist}


def _dft_matrix(n, sign, dtype=np.complex128):
    k = np.arange(n)
    return np.exp(sign * 2j * np.pi * np.outer(k, k) / n).astype(dtype)


def rfft2_forward(x):
    """torch.fft.rfftn(x, dim=(2,3), norm='forward') (shgan.py:313) by explicit DFT matrices:
    X[k1,k2] = 1/(H*W) * sum x[h,w] exp(-2pi i (k1 h/H + k2 w/W)), k2 = 0..W/2."""
    n, c, h, w = x.shape
    cd = np.complex128 if x.dtype == np.float64 else np.complex64
    fw = _dft_matrix(w, -1)[:, :w // 2 + 1]
    fh = _dft_matrix(h, -1)
    y = np.einsum('nchw,wk->nchk', x.astype(np.float64), fw)
    y = np.einsum('lh,nchk->nclk', fh, y) / (h * w)
    return y.astype(cd)


def irfft2_forward_norm(s, dtype):
    """torch.fft.irfftn(s, dim=(2,3), norm='forward') (shgan.py:334) on a NOT necessarily Hermitian
    half spectrum [N,C,r,r/2+1]: unscaled complex inverse DFT along rows (dim 2), then C2R along
    dim 3 in which the imaginary parts of the DC and Nyquist columns are ignored (pocketfft/cuFFT)."""
    n, c, r, hc = s.shape
    wd = 2 * (hc - 1)
    t = np.einsum('lh,nchk->nclk', _dft_matrix(r, +1), s.astype(np.complex128))
    k = np.arange(hc)
    xw = np.arange(wd)
    ang = 2 * np.pi * np.outer(k, xw) / wd
    wgt = np.full(hc, 2.0)
    wgt[0] = 1.0
    wgt[-1] = 1.0
    cosm = np.cos(ang) * wgt[:, None]
    sinm = -np.sin(ang) * wgt[:, None]
    sinm[0] = 0.0
    sinm[-1] = 0.0
    y = np.einsum('nclk,kx->nclx', t.real, cosm) + np.einsum('nclk,kx->nclx', t.imag, sinm)
    return y.astype(dtype)


def shu_forward(sd, x, prefix='encoder.shu', input_res=64, lowest_res=4, freedom=(2, 3), tail_sigma_mult=3,
                gaussian_at_input_res=False, return_stages=False):
    """shgan.py:312-336.  x [N,C,input_res,input_res] -> {r: [N,C,r,r]} for r = lowest_res..input_res."""
    dt = x.dtype
    n, c, h, w = x.shape
    ff = rfft2_forward(x)
    half = h // 2 + 1
    ff = np.concatenate([ff[:, :, half:], ff[:, :, :half]], axis=2)  # :315-317
    t = np.concatenate([ff.real, ff.imag], axis=1).astype(dt)  # :319
    w0 = sd[prefix + '.conv0.weight'].astype(dt)  # [2C,2C,1,1], weight_gain 1 (use_wscale False)
    b0 = sd[prefix + '.conv0.bias'].astype(dt)
    t0 = conv2d(t, w0) + b0.reshape(1, -1, 1, 1)  # :320
    t1 = np.maximum(t0, 0)  # :321
    # heterogeneous_filter :143-160
    dfw = sd[prefix + '.df1.weight'].astype(dt)  # [2C, 2C*fh*fw]
    cw = make_cweight(freedom, (h, w // 2 + 1), dtype=np.float32).astype(dt)
    yk = conv2d(t1, np.ascontiguousarray(dfw.T)[:, :, None, None]).reshape(n, 2 * c, -1, h, w // 2 + 1)
    o = (yk * cw[None, None]).sum(2)
    spec = (o[:, :c] + 1j * o[:, c:])  # :323
    gmaps = gaussian_weight_maps(input_res, lowest_res, tail_sigma_mult, gaussian_at_input_res)
    out = {}
    for r in sorted(gmaps):  # :327-334
        sp = spec[:, :, (input_res // 2 - r // 2):(input_res // 2 + r // 2), 0:(r // 2 + 1)]
        sp = sp * gmaps[r].astype(dt)[None, None]
        sp = np.concatenate([sp[:, :, r - r // 2 - 1:], sp[:, :, :r - r // 2 - 1]], axis=2)
        out[r] = irfft2_forward_norm(sp, dt)
    if return_stages:
        return out, dict(fft_shift=t, conv0=t0, df1=o)
    return out


# ----------------------------------------------------------------------------------------
# encoder / synthesis / generator  (comodgan.py, shgan.py:361-383)
# ----------------------------------------------------------------------------------------

def encoder(sd, img, resolution, shu_channels=32, shu_input_res=64, shu_lowest_res=4, with_shu=True,
            prefix='encoder'):
    """shgan.py:361-383 over comodgan.py:38-64 (encoder_block) and :98-113 (encoder_epilogue),
    has_extra_final_layer=False, mbstd off, dropout in eval mode = identity."""
    log2 = int(np.log2(resolution))
    encode_res = [2 ** i for i in range(log2, 1, -1)]
    feats = {}
    x = None
    for idx, r in enumerate(encode_res[:-1]):
        p = f'{prefix}.b{r}'
        if idx == 0:
            x = conv2d_layer(sd, p + '.fromrgb', img)
        feat = conv2d_layer(sd, p + '.conv0', x)
        x = conv2d_layer(sd, p + '.conv1', feat, down=2)
        feats[r] = feat
    feat = conv2d_layer(sd, f'{prefix}.b4.conv', x)
    feats[4] = feat
    x_global = dense(feat.reshape(feat.shape[0], -1), sd[f'{prefix}.b4.fc.weight'], sd[f'{prefix}.b4.fc.bias'], act=True)
    if with_shu:
        ch = shu_channels
        shu_out = shu_forward(sd, feats[shu_input_res][:, -ch:], prefix=f'{prefix}.shu',
                              input_res=shu_input_res, lowest_res=shu_lowest_res)
        for r, v in shu_out.items():
            f2 = feats[r].copy()
            f2[:, -ch:] = f2[:, -ch:] + v
            feats[r] = f2
    return x_global, feats


def synthesis(sd, x_global, feats, ws, resolution, noise_mode='const', noises=None, prefix='synthesis',
              fused_modconv=True):
    """comodgan.py:396-433 with synthesis_block_first :237-262 and synthesis_block :304-340.
    `noises` (noise_mode='random'): dict layer-prefix -> [N,1,res,res] standard-normal draws."""
    dt = x_global.dtype
    log2 = int(np.log2(resolution))
    block_res = [2 ** i for i in range(2, log2 + 1)]
    noises = noises or {}
    w0 = x_global
    widx = 0
    p = f'{prefix}.b4'
    x = dense(x_global, sd[p + '.fc.weight'], sd[p + '.fc.bias'], act=True)
    x = x.reshape(x.shape[0], -1, 4, 4) + feats[4]
    wl = np.concatenate([ws[:, widx], w0], axis=1)
    x = synthesis_layer(sd, p + '.conv', x, wl, noise_mode=noise_mode, noise=noises.get(p + '.conv'), fused_modconv=fused_modconv)
    wl = np.concatenate([ws[:, widx + 1], w0], axis=1)
    img = torgb_layer(sd, p + '.torgb', x, wl)
    widx += 1
    f = _resample_filter()
    for r in block_res[1:]:
        p = f'{prefix}.b{r}'
        wl = np.concatenate([ws[:, widx], w0], axis=1)
        x = synthesis_layer(sd, p + '.conv0', x, wl, up=2, noise_mode=noise_mode, noise=noises.get(p + '.conv0'), fused_modconv=fused_modconv)
        x = x + feats[r]
        wl = np.concatenate([ws[:, widx + 1], w0], axis=1)
        x = synthesis_layer(sd, p + '.conv1', x, wl, noise_mode=noise_mode, noise=noises.get(p + '.conv1'), fused_modconv=fused_modconv)
        img = upsample2d(img, f)
        wl = np.concatenate([ws[:, widx + 2], w0], axis=1)
        img = img + torgb_layer(sd, p + '.torgb', x, wl)
        widx += 2
    return img.astype(dt)


def generator(sd, x, z, resolution, noise_mode='const', noises=None, with_shu=True, fused_modconv=True,
              return_intermediates=False):
    """comodgan.py:449-481: ws = mapping(z); x_global, feats = encoder(x); img = synthesis(...)."""
    num_ws = {256: 14, 512: 16, 1024: 18}.get(resolution, 2 * int(np.log2(resolution)) - 2)
    ws = mapping(sd, z, num_ws)
    x_global, feats = encoder(sd, x, resolution, with_shu=with_shu)
    img = synthesis(sd, x_global, feats, ws, resolution, noise_mode=noise_mode, noises=noises, fused_modconv=fused_modconv)
    if return_intermediates:
        return img, dict(ws=ws, x_global=x_global, feats=feats)
    return img


# ----------------------------------------------------------------------------------------
# discriminator  (stylegan.py:624-838; comodgan.py:483-485 registers it as comodgan_discriminator)
# ----------------------------------------------------------------------------------------

def minibatch_std(x, group_size=4, num_channels=1):
    """stylegan.py:686-705: std over groups of `group_size` samples, averaged over channels and pixels, appended
    as `num_channels` extra feature maps."""
    n, c, h, w = x.shape
    g = min(group_size, n) if group_size is not None else n
    f = num_channels
    cc = c // f
    y = x.reshape(g, -1, f, cc, h, w)
    y = y - y.mean(axis=0)
    y = np.square(y).mean(axis=0)
    y = np.sqrt(y + x.dtype.type(1e-8))
    y = y.mean(axis=(2, 3, 4))
    y = y.reshape(-1, f, 1, 1)
    y = np.tile(y, (g, 1, h, w))
    return np.concatenate([x, y.astype(x.dtype)], axis=1)


def discriminator(sd, img, resolution, mbstd_group_size=4, mbstd_c_n=1, prefix='', return_intermediates=False):
    """Discriminator.forward (stylegan.py:828-838) with discrim_block.forward (:658-684, reslink=True) and
    discrim_epilogue.forward (:743-755), c_dim == 0."""
    log2 = int(np.log2(resolution))
    res_list = [2 ** i for i in range(log2, 1, -1)]
    p0 = prefix
    x = None
    inter = {}
    for idx, r in enumerate(res_list[:-1]):
        p = f'{p0}b{r}'
        if idx == 0:
            x = conv2d_layer(sd, p + '.fromrgb', img)
        y = conv2d_layer(sd, p + '.skip', x, down=2, act=False, gain=np.sqrt(0.5))
        x = conv2d_layer(sd, p + '.conv0', x)
        x = conv2d_layer(sd, p + '.conv1', x, down=2, gain=np.sqrt(0.5))
        x = y + x
        inter[r // 2] = x
    x = minibatch_std(x, mbstd_group_size, mbstd_c_n) if mbstd_c_n > 0 else x
    x = conv2d_layer(sd, f'{p0}b4.conv', x)
    x = dense(x.reshape(x.shape[0], -1), sd[f'{p0}b4.fc.weight'], sd[f'{p0}b4.fc.bias'], act=True)
    x = dense(x, sd[f'{p0}b4.out.weight'], sd[f'{p0}b4.out.bias'])
    if return_intermediates:
        return x, inter
    return x


def discriminator_state_dict_spec(resolution, ic_n=4, ch_base=32768, ch_max=512, mbstd_c_n=1):
    """Key -> shape of comodgan_discriminator in registration order."""
    log2 = int(np.log2(resolution))
    C = lambda r: channels(r, ch_base, ch_max)
    spec = []
    for i in range(log2, 2, -1):
        r = 2 ** i
        c, cn = C(r), C(r // 2)
        spec.append((f'b{r}.resample_filter', (4, 4)))
        if i == log2:
            spec.append((f'b{r}.fromrgb.weight', (c, ic_n, 1, 1)))
            spec.append((f'b{r}.fromrgb.bias', (c,)))
        spec.append((f'b{r}.conv0.weight', (c, c, 3, 3)))
        spec.append((f'b{r}.conv0.bias', (c,)))
        spec.append((f'b{r}.conv1.weight', (cn, c, 3, 3)))
        spec.append((f'b{r}.conv1.bias', (cn,)))
        spec.append((f'b{r}.conv1.resample_filter', (4, 4)))
        spec.append((f'b{r}.skip.weight', (cn, c, 1, 1)))
        spec.append((f'b{r}.skip.resample_filter', (4, 4)))
    c4 = C(4)
    spec.append(('b4.conv.weight', (c4, c4 + mbstd_c_n, 3, 3)))
    spec.append(('b4.conv.bias', (c4,)))
    spec.append(('b4.fc.weight', (c4, c4 * 16)))
    spec.append(('b4.fc.bias', (c4,)))
    spec.append(('b4.out.weight', (1, c4)))
    spec.append(('b4.out.bias', (1,)))
    return spec


def synthetic_discriminator_state_dict(resolution, seed=0, **kw):
    sd = {}
    f = setup_filter([1, 3, 3, 1])
    for key, shape in discriminator_state_dict_spec(resolution, **kw):
        rng = np.random.Generator(np.random.PCG64(_key_seed(seed, 'D/' + key)))
        if key.endswith('resample_filter'):
            v = f.copy()
        elif key.endswith('.bias'):
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        else:
            v = rng.standard_normal(shape).astype(np.float32)
        sd[key] = np.ascontiguousarray(v)
    return sd


def composite_uint8(x, img):
    """run_generator, lib/experiments/shgan_default.py:257-262."""
    m = x[:, 0:1] + x.dtype.type(0.5)
    out = x[:, 1:4] * m + img * (1 - m)
    out = np.clip(out * x.dtype.type(127.5) + x.dtype.type(127.5), 0, 255)
    return out.astype(np.uint8)  # .to(torch.uint8) truncates toward zero, values are >= 0


# ----------------------------------------------------------------------------------------
# deterministic synthetic weights and inputs shared by fixtures, tests, smoke() and bench.py
# ----------------------------------------------------------------------------------------

def channels(res, ch_base=32768, ch_max=512):
    return min(ch_base // res, ch_max)


def state_dict_spec(resolution, ch_base=32768, ch_max=512, w_dim=512, w0_dim=1024, z_dim=512, num_ws=None,
                    shu_channels=32, ic_n=4):
    """Key -> shape in the reference's registration order."""
    log2 = int(np.log2(resolution))
    if num_ws is None:
        num_ws = {256: 14, 512: 16, 1024: 18}.get(resolution, 2 * log2 - 2)
    C = lambda r: channels(r, ch_base, ch_max)
    spec = []
    spec.append(('mapping.w_avg', (w_dim,)))
    for i in range(8):
        fin = z_dim if i == 0 else w_dim
        spec.append((f'mapping.fc{i}.weight', (w_dim, fin)))
        spec.append((f'mapping.fc{i}.bias', (w_dim,)))
    wl = w_dim + w0_dim

    def syn_layer(p, ci, co, res, has_filter):
        spec.append((p + '.weight', (co, ci, 3, 3)))
        spec.append((p + '.bias', (co,)))
        spec.append((p + '.noise_strength', ()))
        if has_filter:
            spec.append((p + '.resample_filter', (4, 4)))
        spec.append((p + '.noise_const', (res, res)))
        spec.append((p + '.affine.weight', (ci, wl)))
        spec.append((p + '.affine.bias', (ci,)))

    def torgb(p, ci):
        spec.append((p + '.weight', (3, ci, 1, 1)))
        spec.append((p + '.bias', (3,)))
        spec.append((p + '.affine.weight', (ci, wl)))
        spec.append((p + '.affine.bias', (ci,)))
    c4 = C(4)
    spec.append(('synthesis.b4.fc.weight', (c4 * 16, w0_dim)))
    spec.append(('synthesis.b4.fc.bias', (c4 * 16,)))
    syn_layer('synthesis.b4.conv', c4, c4, 4, True)
    torgb('synthesis.b4.torgb', c4)
    for i in range(3, log2 + 1):
        r = 2 ** i
        ci, co = C(r // 2), C(r)
        spec.append((f'synthesis.b{r}.resample_filter', (4, 4)))
        syn_layer(f'synthesis.b{r}.conv0', ci, co, r, True)
        syn_layer(f'synthesis.b{r}.conv1', co, co, r, False)
        torgb(f'synthesis.b{r}.torgb', co)
    for i in range(log2, 2, -1):
        r = 2 ** i
        c, cn = C(r), C(r // 2)
        spec.append((f'encoder.b{r}.resample_filter', (4, 4)))
        if i == log2:
            spec.append((f'encoder.b{r}.fromrgb.weight', (c, ic_n, 1, 1)))
            spec.append((f'encoder.b{r}.fromrgb.bias', (c,)))
        spec.append((f'encoder.b{r}.conv0.weight', (c, c, 3, 3)))
        spec.append((f'encoder.b{r}.conv0.bias', (c,)))
        spec.append((f'encoder.b{r}.conv1.weight', (cn, c, 3, 3)))
        spec.append((f'encoder.b{r}.conv1.bias', (cn,)))
        spec.append((f'encoder.b{r}.conv1.resample_filter', (4, 4)))
    spec.append(('encoder.b4.conv.weight', (c4, c4, 3, 3)))
    spec.append(('encoder.b4.conv.bias', (c4,)))
    spec.append(('encoder.b4.fc.weight', (w0_dim, c4 * 16)))
    spec.append(('encoder.b4.fc.bias', (w0_dim,)))
    spec.append(('encoder.shu.conv0.weight', (2 * shu_channels, 2 * shu_channels, 1, 1)))
    spec.append(('encoder.shu.conv0.bias', (2 * shu_channels,)))
    spec.append(('encoder.shu.df1.weight', (2 * shu_channels, 2 * shu_channels * 6)))
    return spec


def _key_seed(seed, key):
    h = 1469598103934665603
    for ch in (str(seed) + '/' + key).encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def synthetic_state_dict(resolution, seed=0, **kw):
    """Deterministic, constructor-order-independent random weights (numpy PCG64 keyed on the
    state_dict key) with the reference's init statistics, except that biases, noise strengths
    and df1 are randomised so that every term of the path is exercised."""
    sd = {}
    f = setup_filter([1, 3, 3, 1])
    for key, shape in state_dict_spec(resolution, **kw):
        rng = np.random.Generator(np.random.PCG64(_key_seed(seed, key)))
        if key.endswith('resample_filter'):
            v = f.copy()
        elif key == 'mapping.w_avg':
            v = np.zeros(shape, np.float32)
        elif key.endswith('noise_strength'):
            v = np.float32(rng.standard_normal() * 0.1).reshape(())
        elif key.endswith('affine.bias'):
            v = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif key.endswith('.bias'):
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif key.startswith('mapping.fc') and key.endswith('weight'):
            v = (rng.standard_normal(shape) / 0.01).astype(np.float32)  # dense: randn / lr_multi
        elif key == 'encoder.shu.df1.weight':
            v = (1 / 64 + (0.5 / 64) * rng.standard_normal(shape)).astype(np.float32)
        elif key == 'encoder.shu.conv0.weight':
            v = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(np.float32)
        else:
            v = rng.standard_normal(shape).astype(np.float32)
        sd[key] = np.ascontiguousarray(v)
    return sd


def synthetic_inputs(batch, resolution, seed=0, z_dim=512):
    """x = cat([mask-0.5, img*mask]) exactly as the eval loop builds it (shgan_default.py:269-276),
    with a cheap deterministic free-form-like mask (rectangles + thick strokes; 1 = keep)."""
    rng = np.random.Generator(np.random.PCG64(_key_seed(seed, f'inputs{batch}x{resolution}')))
    img = np.clip(rng.standard_normal((batch, 3, resolution, resolution)), -1, 1).astype(np.float32)
    mask = np.ones((batch, 1, resolution, resolution), np.float32)
    s = resolution
    for b in range(batch):
        for _ in range(int(rng.integers(1, 5))):
            w, h = int(rng.integers(s // 8, s // 2)), int(rng.integers(s // 8, s // 2))
            x0, y0 = int(rng.integers(0, s - w)), int(rng.integers(0, s - h))
            mask[b, 0, y0:y0 + h, x0:x0 + w] = 0
        for _ in range(int(rng.integers(1, 4))):
            px, py = float(rng.integers(0, s)), float(rng.integers(0, s))
            width = int(rng.integers(max(2, s // 40), max(3, s // 10)))
            for _ in range(int(rng.integers(2, 8))):
                ang, ln = rng.uniform(0, 2 * math.pi), rng.uniform(s / 16, s / 3)
                qx, qy = np.clip(px + ln * math.cos(ang), 0, s - 1), np.clip(py + ln * math.sin(ang), 0, s - 1)
                steps = int(max(abs(qx - px), abs(qy - py))) + 1
                for t in np.linspace(0, 1, steps):
                    cx, cy = int(px + (qx - px) * t), int(py + (qy - py) * t)
                    mask[b, 0, max(cy - width // 2, 0):cy + width // 2 + 1, max(cx - width // 2, 0):cx + width // 2 + 1] = 0
                px, py = qx, qy
    x = np.concatenate([mask - 0.5, img * mask], axis=1).astype(np.float32)
    z = rng.standard_normal((batch, z_dim)).astype(np.float32)
    return x, z
