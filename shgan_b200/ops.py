"""Operator-level API with the reference's names, argument meaning and error behaviour, executed by the
sm_100a kernels of this package (NCHW fp32 CUDA tensors in and out, like the reference ops):

    setup_filter / upfirdn2d / upsample2d / downsample2d / filter2d   <- stylegan_utils/upfirdn2d.py:66,198,245,279,316
    conv2d_resample                                                    <- stylegan_utils/conv2d_resample.py:57
    modulated_conv2d                                                   <- stylegan.py:103-193
    lrelu_agc                                                          <- common/utils.py:117-146

These entry points convert to/from the split-plane layout at their boundary; the fused generator
(`engine.GeneratorEngine`) calls the same kernels without those conversions.  There is no CPU or PyTorch
fallback: non-CUDA inputs raise.
"""
import math

import torch

from . import kernels as K
from . import packing as P
from .kernels import Planes

setup_filter = P.setup_filter


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    return tuple(padding)


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    return int(f.shape[-1]), int(f.shape[0])


def _require_cuda(x, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError(f'{what}: shgan_b200 has no CPU path; expected a CUDA tensor')


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, upsample, filter and downsample a batch of 2D images (upfirdn2d.py:198-239).  `impl` is accepted for
    signature compatibility; the sm_100a kernel is the only implementation."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    _require_cuda(x, 'upfirdn2d')
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    assert f.dtype == torch.float32 and f.ndim in [1, 2]
    f = f.to(x.device)
    xin = x.contiguous().float()
    if f.ndim == 2:
        y = K.upfirdn2d_fwd(xin, f.contiguous(), upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    else:  # separable: one pass per axis (upfirdn2d.py:160-162)
        y = K.upfirdn2d_fwd(xin, f.unsqueeze(0).contiguous(), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, math.sqrt(gain))
        y = K.upfirdn2d_fwd(y, f.unsqueeze(1).contiguous(), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, math.sqrt(gain))
    return y.to(x.dtype)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def lrelu_agc(x, alpha=0.2, gain=math.sqrt(2.0), clamp=256, extra_gain=1):
    """common/utils.py:135-143 (tiny tensors only reach this op-level helper; the hot path fuses it)."""
    x = torch.nn.functional.leaky_relu(x, negative_slope=alpha)
    g = gain * extra_gain
    if g != 1:
        x = x * g
    if clamp is not None:
        x = x.clamp(-clamp * extra_gain, clamp * extra_gain)
    return x


# ---- convolution helpers ------------------------------------------------------------------------------
def _pad64(c):
    return (c + 63) // 64 * 64


def _to_planes_padded(x, scale=None):
    """NCHW fp32 -> Planes with the channel count zero-padded to a multiple of 64."""
    n, c, h, w = x.shape
    cp = _pad64(c)
    out = Planes.empty(n, h, w, cp, x.device)
    K.nchw_to_planes(x.contiguous().float(), scale=scale, out=out, c_off=0)
    return out


def _pack_padded(w):
    co, ci, kh, kw = w.shape
    wp = torch.zeros((_pad64(co), _pad64(ci), kh, kw), dtype=torch.float32, device=w.device)
    wp[:co, :ci] = w
    return P.pack_conv_weight(wp)


def _conv_planes(xp, w, taps_fn, oh, ow, epi_kwargs, passes, impl):
    """xp: Planes (padded channels), w [Co,Ci,kh,kw].  Returns NCHW fp32 [N,Co,oh,ow]."""
    co = w.shape[0]
    cop = _pad64(co)
    n = xp.shape[0]
    wh, wl = _pack_padded(w)
    y = torch.empty((n, oh, ow, cop), dtype=torch.float32, device=w.device)
    epi = K.make_epilogue(out_f32=y, **epi_kwargs)
    K.conv_igemm([xp], wh, wl, taps_fn, oh, ow, epi=epi, passes=passes, impl=impl)
    return K.nhwc_to_nchw_f32(y)[:, :co].contiguous()


def _pad_vec(v, cp):
    if v is None:
        return None
    out = torch.zeros(v.shape[:-1] + (cp,), dtype=torch.float32, device=v.device)
    out[..., :v.shape[-1]] = v
    return out.contiguous()


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False,
                    _scale=None, _epi=None, passes=3, impl=0):
    """2D convolution with optional up/downsampling (conv2d_resample.py:57-154), same decomposition as the
    reference: blur -> strided conv for down=2, transposed conv -> blur for up=2, 1x1 fast paths, plain conv otherwise.
    `_scale` ([N,Ci] per-sample input modulation) and `_epi` (epilogue terms applied after the conv/blur) are
    used by modulated_conv2d."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    _require_cuda(x, 'conv2d_resample')
    if groups != 1:
        raise NotImplementedError('grouped convolution is not on the SH-GAN generator path (modulation is applied to the '
                                  'activations instead of materialising per-sample weights)')
    co, ci, kh, kw = [int(s) for s in w.shape]
    assert x.shape[1] == ci
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    n, _, h, wd = x.shape
    epi = dict(_epi or {})
    cop = _pad64(co)
    for k in ('dcoef', 'bias'):
        if epi.get(k) is not None:
            epi[k] = _pad_vec(epi[k], cop)
    wc = w.float() if flip_weight else w.float().flip([2, 3])   # correlation weights
    scale = _scale.contiguous().float() if _scale is not None else None   # [N,Ci], indexed with x's own channel count

    if kw == 1 and kh == 1 and down > 1 and up == 1:      # :105-108
        xd = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        xp = _to_planes_padded(xd, scale)
        return _conv_planes(xp, wc, P.taps_plain(1, 1), xd.shape[2], xd.shape[3], epi, passes, impl)
    if kw == 1 and kh == 1 and up > 1 and down == 1:      # :111-114
        if _epi:
            raise NotImplementedError('fused epilogue after a 1x1 up-sampling conv')
        xp = _to_planes_padded(x, scale)
        y = _conv_planes(xp, wc, P.taps_plain(1, 1), h, wd, {}, passes, impl)
        return upfirdn2d(y, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                               # :117-120
        if not (down == 2 and kh == 3 and kw == 3 and fw == 4 and fh == 4 and f.ndim == 2):
            raise NotImplementedError('down-sampling conv supports down=2, 3x3 weights and a 4x4 filter')
        xp = _to_planes_padded(x, scale)
        oh_b, ow_b = h + py0 + py1 - 3, wd + px0 + px1 - 3
        ph, pw = (oh_b + 1) // 2, (ow_b + 1) // 2
        par = Planes.empty(4 * n, ph, pw, xp.shape[3], x.device)
        fa = f.float() if flip_filter else f.float().flip([0, 1])
        K.fir_nhwc(xp, fa.contiguous(), 1.0, (px0, px1, py0, py1), K.make_epilogue(out=par), parity_split=True)
        srcs = [Planes(par.hi[q * n:(q + 1) * n], par.lo[q * n:(q + 1) * n]) for q in range(4)]
        oh, ow = (oh_b - 3) // 2 + 1, (ow_b - 3) // 2 + 1
        wh, wl = _pack_padded(wc)
        y = torch.empty((n, oh, ow, cop), dtype=torch.float32, device=x.device)
        K.conv_igemm(srcs, wh, wl, P.taps_down2(3), oh, ow, epi=K.make_epilogue(out_f32=y, **epi), passes=passes, impl=impl)
        return K.nhwc_to_nchw_f32(y)[:, :co].contiguous()
    if up > 1:                                             # :123-142
        if not (up == 2 and down == 1 and kh == 3 and kw == 3 and fw == 4 and fh == 4 and f.ndim == 2):
            raise NotImplementedError('up-sampling conv supports up=2, 3x3 weights and a 4x4 filter')
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0); pyt = max(min(-py0, -py1), 0)
        if pxt != 0 or pyt != 0:
            raise NotImplementedError('cropped transposed convolution')
        # conv_transpose2d is called with flip_weight=(not flip_weight): true convolution when flip_weight is False
        wt = w.float() if not flip_weight else w.float().flip([2, 3])
        xp = _to_planes_padded(x, scale)
        wh, wl = _pack_padded(wt)
        z = torch.empty((n, 2 * h + 1, 2 * wd + 1, cop), dtype=torch.float32, device=x.device)
        for py in range(2):
            for pxx in range(2):
                K.conv_igemm([xp], wh, wl, P.taps_up2(py, pxx), P.up2_pass_size(h, py), P.up2_pass_size(wd, pxx),
                             raw=(z, 2, 2, py, pxx), passes=passes, impl=impl)
        oh, ow = 2 * h + 1 + py0 + py1 - 3, 2 * wd + 1 + px0 + px1 - 3
        y = torch.empty((n, oh, ow, cop), dtype=torch.float32, device=x.device)
        fa = f.float() if flip_filter else f.float().flip([0, 1])
        K.fir_nhwc(z, fa.contiguous(), float(up ** 2), (px0, px1, py0, py1), K.make_epilogue(out_f32=y, **epi))
        return K.nhwc_to_nchw_f32(y)[:, :co].contiguous()
    if up == 1 and down == 1:                              # :145-147
        if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0 and px0 == kw // 2 and py0 == kh // 2 and kh * kw <= K._lib.SHGAN_MAX_TAPS:
            xp = _to_planes_padded(x, scale)
            return _conv_planes(xp, wc, P.taps_plain(kh, kw), h, wd, epi, passes, impl)
    raise NotImplementedError('this conv2d_resample configuration is not reached by the SH-GAN generator')


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True, passes=3, impl=0):
    """stylegan.py:103-193.  Modulation is applied to the activations, demodulation in the conv epilogue
    (mathematically the reference's two branches are identical, so `fused_modconv` only exists for signature
    compatibility); weights are never materialised per sample."""
    _require_cuda(x, 'modulated_conv2d')
    n = int(x.shape[0])
    co, ci, kh, kw = [int(s) for s in weight.shape]
    assert x.shape[1] == ci and tuple(styles.shape) == (n, ci)
    styles = styles.contiguous().float()
    s_hat = torch.empty_like(styles)
    dcoef = None
    w = weight.float()
    if demodulate:
        w, wsq = P.demod_weight(weight)
        dcoef = torch.empty((n, co), dtype=torch.float32, device=x.device)
        K.style_prep(styles, wsq, s_hat, dcoef, True)
    else:
        K.style_prep(styles, None, s_hat, None, False, 1.0)
    epi = dict(dcoef=dcoef)
    keep = []
    if noise is not None:
        nz = noise.float().contiguous()
        oh = x.shape[2] * up // down
        ow = x.shape[3] * up // down
        if nz.numel() == oh * ow:
            epi.update(noise=nz, noise_sn=0)
        else:
            nz = nz.expand(n, 1, oh, ow).contiguous()
            epi.update(noise=nz, noise_sn=oh * ow)
        one = torch.ones([], dtype=torch.float32, device=x.device)
        epi.update(noise_strength=one)
        keep += [nz, one]
    return conv2d_resample(x, w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight,
                           _scale=s_hat, _epi=epi, passes=passes, impl=impl)
