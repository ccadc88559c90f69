"""Host-side, one-time preparation of kernel operands: fp16 hi/lo splitting, tap-major weight packing, the
tap tables that turn the reference's resampled convolutions into shgan_conv_igemm descriptors, and the
SHU constants.  Pure torch/numpy tensor algebra with no device dependence, so the tables are unit-tested
on CPU against the oracle (tests/test_host_logic.py).
"""
import math

import numpy as np
import torch


def split_f16(t):
    """value -> (hi, lo) fp16 with hi + lo == value to ~22 significant bits."""
    t = t.to(torch.float32)
    hi = t.to(torch.float16)
    lo = (t - hi.to(torch.float32)).to(torch.float16)
    return hi.contiguous(), lo.contiguous()


def pack_conv_weight(w):
    """[Co,Ci,kh,kw] fp32 -> (w_hi, w_lo) fp16 [kh*kw, Co, Ci]; tap index = ky*kw + kx."""
    co, ci, kh, kw = w.shape
    wt = w.detach().to(torch.float32).permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    return split_f16(wt)


UP2_TAP_ORDER = (3, 0, 1, 4, 5, 2, 6, 7, 8)


def pack_up2_weight(w):
    """[Co,Ci,3,3] fp32 -> (w_hi, w_lo) fp16 [Co/64, 9, 64, Ci] for shgan_conv_up2: per block of 64 output channels the
    nine taps in the order its stacked-N MMAs read them (include/shgan_b200.h): operand shift (0,0) feeds parities
    (1,0) (0,0) (0,1) (1,1) with taps 3,0,1,4; shift (0,-1) feeds (1,0) (0,0) with 5,2; shift (-1,0) feeds (0,0) (0,1)
    with 6,7; shift (-1,-1) feeds (0,0) with 8 (tap = ky*3 + kx of the un-flipped weight, cf. taps_up2)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3 and co % 64 == 0
    wt = w.detach().to(torch.float32).permute(2, 3, 0, 1).reshape(9, co // 64, 64, ci)      # [tap, blk, o, i]
    wt = wt[list(UP2_TAP_ORDER)].permute(1, 0, 2, 3).contiguous()                            # [blk, slot, o, i]
    return split_f16(wt)


def separable_taps(f):
    """4x4 filter (as applied) -> (fy, fx) python lists with outer(fy, fx) == f, or None if f is not rank 1."""
    f = torch.as_tensor(f, dtype=torch.float64).cpu()
    if f.ndim != 2 or tuple(f.shape) != (4, 4) or float(f.sum()) == 0.0:
        return None
    fy, fx = f.sum(dim=1), f.sum(dim=0) / f.sum()
    if float((torch.outer(fy, fx) - f).abs().max()) > 1e-7 * float(f.abs().max()):
        return None
    return [float(v) for v in fy], [float(v) for v in fx]


def demod_weight(w):
    """Weight pre-normalisation of modulated_conv2d (stylegan.py:146) and the per-(o,i) energy table
    wsq[o,i] = sum_k w_hat[o,i,k]^2 from which the demodulation coefficients are formed (stylegan.py:155)."""
    w = w.detach().to(torch.float32)
    w_hat = w * w.square().mean(dim=[1, 2, 3], keepdim=True).rsqrt()
    wsq = w_hat.square().sum(dim=[2, 3]).contiguous()
    return w_hat, wsq


# ---- tap tables: lists of (source, dy, dx, weight_tap) --------------------------------------------------
def taps_plain(kh, kw):
    """stride-1 correlation with padding k//2 (conv2d_resample.py:145-147): out[y,x] += w[ky,kx]*in[y+ky-ph, x+kx-pw]."""
    return [(0, ky - kh // 2, kx - kw // 2, ky * kw + kx) for ky in range(kh) for kx in range(kw)]


def taps_down2(k=3):
    """stride-2, pad-0 correlation over a blurred image that the FIR kernel stored de-interleaved into four
    parity planes P[(y&1)*2+(x&1)][y>>1, x>>1] (conv2d_resample.py:117-120): Bl[2oy+ky, 2ox+kx]."""
    return [((ky & 1) * 2 + (kx & 1), ky >> 1, kx >> 1, ky * k + kx) for ky in range(k) for kx in range(k)]


def taps_up2(py, px, k=3):
    """Parity pass (py,px) of the stride-2 transposed convolution z[2i+ky, 2j+kx] += x[i,j]*w[ky,kx]
    (conv2d_resample.py:123-137 with the un-flipped weights selected by flip_weight=False): the outputs
    z[2a+py, 2b+px] only receive the taps with ky = py, kx = px (mod 2), read at x[a+(py-ky)/2, b+(px-kx)/2]."""
    return [(0, (py - ky) // 2, (px - kx) // 2, ky * k + kx)
            for ky in range(k) if ky % 2 == py for kx in range(k) if kx % 2 == px]


def up2_pass_size(h, p):
    """Number of outputs of parity p along an axis of input length h (z has 2h+1 samples)."""
    return h + 1 if p == 0 else h


# ---- activation strings (common/utils.py:40-87, 117-146) --------------------------------------------------
def parse_activation(s):
    """'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)' -> dict(alpha, gain, clamp) ; None -> None."""
    if s is None:
        return None
    s = s.strip()
    if not s.startswith('lrelu_agc'):
        raise NotImplementedError(f'activation {s!r} is not on the SH-GAN generator path')
    out = dict(alpha=0.1, gain=1.0, clamp=None)
    body = s[len('lrelu_agc'):].strip()
    if body.startswith('('):
        for item in body.strip('()').split(','):
            if not item.strip():
                continue
            k, v = [t.strip() for t in item.split('=')]
            if v == 'sqrt_2':
                out[k] = math.sqrt(2.0)
            elif v == 'None':
                out[k] = None
            else:
                out[k] = float(v)
    return out


# ---- filters (stylegan_utils/upfirdn2d.py:66-92) -----------------------------------------------------------
def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


# ---- SHU constants (shgan.py:70-121, 280-310) ------------------------------------------------------------------
def make_cweight(half_size, half_sample):
    """Blend weights of the heterogeneous filter for type='piecewise_linear', oddeven_aligned=True
    (shgan.py:70-121).  The reference bilinearly samples (align_corners=True) a one-hot lattice of
    h0 x w0 anchors, which is the product of two hat functions of the lattice coordinates."""
    h0, w0 = half_size
    hs, ws = half_sample
    if hs % 2 == 0:
        hg = np.array([-1 + i / hs * 2 for i in range(hs + 1)])[1:]
    else:
        hg = np.array([-1 + i / (hs - 1) * 2 for i in range(hs)])
    wg = np.array([i / (ws - 1) for i in range(ws)])
    hg = hg.astype(np.float32).astype(np.float64)
    wg = wg.astype(np.float32).astype(np.float64)
    v = (hg + 1) / 2 * (h0 - 1)
    u = wg * (w0 - 1)
    hat = lambda t: np.maximum(0.0, 1.0 - np.abs(t))
    cw = np.zeros((h0 * w0, hs, ws), np.float64)
    for r in range(h0):
        for c in range(w0):
            cw[r * w0 + c] = hat(v - r)[:, None] * hat(u - c)[None, :]
    return torch.from_numpy(cw.astype(np.float32))


def gaussian_band_masks(input_res, lowest_res, tail_sigma_mult=3, gaussian_at_input_res=False):
    """Band-splitting masks of SHU.__init__ (shgan.py:280-310): {r: float32 [r, r/2+1]}.  Each band is an
    isotropic Gaussian centred on the (shifted) DC bin with sigma = (r/2)/tail_sigma_mult, and every larger
    band has the next smaller band's Gaussian subtracted from its centre window."""
    reslist = [2 ** i for i in range(int(np.log2(lowest_res)), int(np.log2(input_res)) + 1)]
    rev = reslist[::-1]
    maps = {}

    def gauss(r):
        sigma = (r // 2) / tail_sigma_mult
        ci, cj = r // 2 - 1, 0
        hh = np.arange(r, dtype=np.float64)[:, None]
        ww = np.arange(r // 2 + 1, dtype=np.float64)[None, :]
        g = np.zeros((r, r // 2 + 1), np.float64)
        sr = int(3 * sigma + 1)   # the reference only evaluates a +-3 sigma window (gaussian_heatmap_2d speed-up)
        h0_, h1_ = max(min(ci - sr, r), 0), max(min(ci + sr, r), 0)
        w0_, w1_ = max(min(cj - sr, r // 2 + 1), 0), max(min(cj + sr, r // 2 + 1), 0)
        e = np.exp(-0.5 * (((hh - ci) ** 2) + ((ww - cj) ** 2)) / (sigma ** 2))
        g[h0_:h1_, w0_:w1_] = e[h0_:h1_, w0_:w1_]
        return g
    for idx, r in enumerate(rev):
        if idx != 0:
            maps[r] = gauss(r)
            rp = rev[idx - 1]
            maps[rp][(rp // 2 - r // 2):(rp // 2 + r // 2), 0:(r // 2 + 1)] -= maps[r]
        elif gaussian_at_input_res:
            maps[r] = gauss(r)
        else:
            maps[r] = np.ones((r, r // 2 + 1), np.float64)
    return {r: torch.from_numpy(maps[r].astype(np.float32)) for r in reslist}
