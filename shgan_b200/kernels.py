"""Thin torch-tensor wrappers over the C ABI (include/shgan_b200.h).

PyTorch is used only for device memory and streams: every function takes CUDA tensors, passes their raw
device pointers plus the current torch stream to the library, and returns the caller-visible outputs.
Nothing here computes on the CPU and there is no fallback implementation.
"""
import ctypes as C
import functools
import math
import os

import torch

from . import _lib
from ._lib import ConvDesc, Epilogue, Up2Desc

SQRT2 = math.sqrt(2.0)
ACC_COMP = float(os.environ.get('SHGAN_ACC_COMP', '0'))   # development override of shgan_conv_desc::acc_comp


def _stream():
    # called inside `_on_tensor_device`, so the current device is the tensors' device
    return torch.cuda.current_stream().cuda_stream


def _find_devices(obj, found):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.add(obj.device.index)
    elif isinstance(obj, Planes):
        _find_devices(obj.hi, found)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _find_devices(o, found)
    elif isinstance(obj, dict):
        for o in obj.values():
            _find_devices(o, found)


def _on_tensor_device(fn):
    """Every launch goes to the device that owns the tensors, on that device's current torch stream -- not to whatever
    device happens to be current (the reference eval puts rank r's model on cuda:r without torch.cuda.set_device,
    lib/experiments/shgan_default.py:164; its own op runs under an OptionalCUDAGuard, upfirdn2d.cpp:31).  Tensors on
    two different devices in one call are an error."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        found = set()
        _find_devices(args, found)
        _find_devices(kwargs, found)
        if len(found) > 1:
            raise RuntimeError(f'{fn.__name__}: tensors live on different CUDA devices {sorted(found)}')
        if not found or next(iter(found)) == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(found))):
            return fn(*args, **kwargs)
    return wrapped


def _p(t):
    return None if t is None else t.data_ptr()


def _f32c(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f'{name} must be a contiguous float32 CUDA tensor')
    return t


class Planes:
    """Split-plane NHWC activation: value = hi + lo, both fp16 [N,H,W,C] (include/shgan_b200.h)."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(n, h, w, c, device):
        # zero-filled once: rows/columns that kernels never write (parity padding) must not hold NaN patterns
        return Planes(torch.zeros((n, h, w, c), dtype=torch.float16, device=device),
                      torch.zeros((n, h, w, c), dtype=torch.float16, device=device))

    @property
    def shape(self):
        return tuple(self.hi.shape)

    def float(self):
        """fp32 NHWC view of the value (test helper; not used on the hot path)."""
        return self.hi.float() + self.lo.float()


def make_epilogue(dcoef=None, wgain=1.0, noise=None, noise_sn=0, noise_strength=None, bias=None, act=False,
                  act_alpha=0.2, act_gain=1.0, act_clamp=-1.0, skip=None, next_scale=None, rgb_w=None,
                  rgb_style=None, rgb_out=None, out=None, out_f32=None):
    e = Epilogue()
    e.dcoef = _p(dcoef); e.wgain = wgain
    e.noise = _p(noise); e.noise_sn = noise_sn; e.noise_strength = _p(noise_strength)
    e.bias = _p(bias); e.act = 1 if act else 0
    e.act_alpha = act_alpha; e.act_gain = act_gain; e.act_clamp = act_clamp
    e.skip_hi = _p(skip.hi) if skip is not None else None
    e.skip_lo = _p(skip.lo) if skip is not None else None
    e.next_scale = _p(next_scale); e.rgb_w = _p(rgb_w); e.rgb_style = _p(rgb_style); e.rgb_out = _p(rgb_out)
    e.out_hi = _p(out.hi) if out is not None else None
    e.out_lo = _p(out.lo) if out is not None else None
    e.out_f32 = _p(out_f32)
    return e


# ---- upfirdn2d -----------------------------------------------------------------------------------
@_on_tensor_device
def upfirdn2d_fwd(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """Same contract as the reference pybind op (upfirdn2d.cpp:16): allocates and returns y."""
    _f32c(x, 'x'); _f32c(f, 'f')
    n, c, h, w = x.shape
    fh, fw = f.shape
    ow = (w * upx + padx0 + padx1 - fw + downx) // downx
    oh = (h * upy + pady0 + pady1 - fh + downy) // downy
    if ow < 1 or oh < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')
    y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.shgan_upfirdn2d_fwd(_p(x), _p(f), _p(y), n, c, h, w, fh, fw, upx, upy, downx, downy,
                                      padx0, padx1, pady0, pady1, 1 if flip else 0, float(gain), _stream()),
               'shgan_upfirdn2d_fwd')
    return y


# ---- layout ----------------------------------------------------------------------------------------
@_on_tensor_device
def nchw_to_planes(x, add=None, scale=None, out=None, c_off=0):
    _f32c(x, 'x')
    n, c, h, w = x.shape
    if out is None:
        out = Planes.empty(n, h, w, c, x.device)
    c_tot = out.shape[3]
    lib = _lib.load()
    _lib.check(lib.shgan_nchw_to_planes(_p(x), _p(add.hi) if add is not None else None,
                                       _p(add.lo) if add is not None else None, _p(scale), _p(out.hi), _p(out.lo),
                                       n, c, h, w, c_off, c_tot, _stream()), 'shgan_nchw_to_planes')
    return out


@_on_tensor_device
def planes_to_nchw(p, c_off=0, c=None, out=None):
    n, h, w, c_tot = p.shape
    c = c_tot - c_off if c is None else c
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=p.hi.device)
    lib = _lib.load()
    _lib.check(lib.shgan_planes_to_nchw(_p(p.hi), _p(p.lo), _p(out), n, c, h, w, c_off, c_tot, _stream()),
               'shgan_planes_to_nchw')
    return out


@_on_tensor_device
def planes_add_nchw(p, x, c_off):
    _f32c(x, 'x')
    n, c, h, w = x.shape
    lib = _lib.load()
    _lib.check(lib.shgan_planes_add_nchw(_p(p.hi), _p(p.lo), _p(x), n, c, h, w, c_off, p.shape[3], _stream()),
               'shgan_planes_add_nchw')
    return p


@_on_tensor_device
def planes_add_nchw_multi(planes_list, xs, c_off_list):
    """planes_list[k][..., c_off:c_off+C] += xs[k] (NCHW fp32 [N,C,h,w]) for all k in one launch."""
    b = _lib.AddBatch()
    b.num = len(xs)
    n, c = xs[0].shape[0], xs[0].shape[1]
    for k, (p, x, co) in enumerate(zip(planes_list, xs, c_off_list)):
        _f32c(x, 'x')
        assert x.shape[0] == n and x.shape[1] == c and tuple(p.shape[1:3]) == tuple(x.shape[2:])
        b.hi[k] = _p(p.hi); b.lo[k] = _p(p.lo); b.x[k] = _p(x)
        b.hw[k] = x.shape[2] * x.shape[3]; b.c_off[k] = co; b.c_tot[k] = p.shape[3]
    lib = _lib.load()
    _lib.check(lib.shgan_planes_add_nchw_multi(C.byref(b), n, c, _stream()), 'shgan_planes_add_nchw_multi')


@_on_tensor_device
def nhwc_to_nchw_f32(x):
    _f32c(x, 'x')
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.shgan_nhwc_to_nchw_f32(_p(x), _p(y), n, c, h, w, _stream()), 'shgan_nhwc_to_nchw_f32')
    return y


# ---- convolution -------------------------------------------------------------------------------------
@_on_tensor_device
def conv_igemm(srcs, w_hi, w_lo, taps, oh, ow, epi=None, raw=None, block_n=0, passes=3, impl=0, acc_comp=None, splitk=None):
    """srcs: list of Planes [N,Hs,Ws,C]; w_hi/w_lo: fp16 [w_taps, Co, C]; taps: list of (src, dy, dx, w_tap).
    epi: Epilogue (ACT mode)  or  raw = (z fp32 [N,ZH,ZW,Co], zsy, zsx, zoy, zox) (RAW mode).
    splitk (ACT mode): an fp32 scratch tensor; layers with too few output tiles to fill the GPU (4x4, 8x8) are then split along K
    over several CTAs per tile (the library decides; the scratch must hold at least two partial outputs)."""
    d = ConvDesc()
    d.num_src = len(srcs)
    n, _, _, c = srcs[0].shape
    for i, s in enumerate(srcs):
        d.src_hi[i] = _p(s.hi); d.src_lo[i] = _p(s.lo)
        d.src_h[i] = s.shape[1]; d.src_w[i] = s.shape[2]
    d.N = n; d.C = c; d.Co = w_hi.shape[1]
    d.w_hi = _p(w_hi); d.w_lo = _p(w_lo); d.w_taps = w_hi.shape[0]
    d.ntaps = len(taps)
    for i, (s, dy, dx, wt) in enumerate(taps):
        d.tap_src[i] = s; d.tap_dy[i] = dy; d.tap_dx[i] = dx; d.tap_w[i] = wt
    d.OH = oh; d.OW = ow
    if raw is not None:
        z, zsy, zsx, zoy, zox = raw
        d.mode = 1; d.z = _p(z); d.ZH = z.shape[1]; d.ZW = z.shape[2]
        d.zsy = zsy; d.zsx = zsx; d.zoy = zoy; d.zox = zox
    else:
        d.mode = 0
        d.epi = epi
        if splitk is not None:
            per = n * oh * ow * d.Co
            nbuf = min(int(splitk.numel() // per), 64)
            if nbuf >= 2:
                d.z = _p(splitk); d.ZH = nbuf; d.ZW = per
    d.block_n = block_n; d.passes = passes; d.impl = impl
    d.acc_comp = ACC_COMP if acc_comp is None else acc_comp    # 0 = library default, < 0 = off (include/shgan_b200.h)
    if impl == 1:       # fp32 FMA cross-check: lives in the test-only library, not in libshgan_b200.so
        chk = _lib.load_check()
        rc = chk.shgan_check_conv_igemm(C.byref(d), _stream())
        if rc != 0:
            raise RuntimeError(f'shgan_check_conv_igemm failed (code {rc}): {chk.shgan_last_error().decode()}')
        return
    lib = _lib.load()
    _lib.check(lib.shgan_conv_igemm(C.byref(d), _stream()), 'shgan_conv_igemm')


UP2_NARROW = 0x100      # SHGAN_UP2_NARROW


@_on_tensor_device
def conv_up2(src, w_hi, w_lo, fx, fy, gain, epi, passes=3, acc_comp=None, narrow=False, cluster=None):
    """Fused up-sampling convolution (shgan_conv_up2): src Planes [N,H,W,C]; w_hi/w_lo fp16 [Co/64, 9, 64, C] from
    packing.pack_up2_weight; fx/fy: the separable blur taps as applied (4 floats each); epi: Epilogue at [N,2H,2W,Co].
    narrow=True forces the 8-warp epilogue instance, cluster=True / False forces / forbids the two-CTA weight-sharing clusters
    (tests / profiling; None = the library's choice)."""
    d = Up2Desc()
    n, h, w, c = src.shape
    d.src_hi = _p(src.hi); d.src_lo = _p(src.lo)
    d.N = n; d.H = h; d.W = w; d.C = c; d.Co = w_hi.shape[0] * 64
    d.w_hi = _p(w_hi); d.w_lo = _p(w_lo)
    for i in range(4):
        d.fx[i] = float(fx[i]); d.fy[i] = float(fy[i])
    d.gain = float(gain)
    d.epi = epi
    d.passes = passes | (UP2_NARROW if narrow else 0) | (0 if cluster is None else (0x200 if cluster else 0x400))
    d.acc_comp = ACC_COMP if acc_comp is None else acc_comp
    lib = _lib.load()
    _lib.check(lib.shgan_conv_up2(C.byref(d), _stream()), 'shgan_conv_up2')


def conv_num_nblocks(co, block_n=0):
    return _lib.load().shgan_conv_num_nblocks(co, block_n)


FIR_RANK1 = 0x100


FIR_TWO_PHASE = 0x200   # SHGAN_FIR_TWO_PHASE


@_on_tensor_device
def fir_nhwc(src, f, gain, pads, epi, parity_split=False, rank1=False, two_phase=False):
    """src: Planes or fp32 NHWC tensor; pads = (pad_x0, pad_x1, pad_y0, pad_y1); f: fp32 [4,4] (as applied).
    rank1=True: the caller has checked (on the host, once) that f is separable -> no fallback launch (SHGAN_FIR_RANK1);
    with planes in and an identity epilogue that is the row-walking kernel unless two_phase=True (tests, profiling)."""
    lib = _lib.load()
    if isinstance(src, Planes):
        n, ih, iw, c = src.shape
        args = (None, _p(src.hi), _p(src.lo))
    else:
        _f32c(src, 'src')
        n, ih, iw, c = src.shape
        args = (_p(src), None, None)
    _lib.check(lib.shgan_fir_nhwc(*args, _p(f), f.shape[0], f.shape[1], float(gain), n, c, ih, iw,
                                 pads[0], pads[1], pads[2], pads[3], C.byref(epi), int(parity_split) | (FIR_RANK1 if rank1 else 0) | (FIR_TWO_PHASE if two_phase else 0), _stream()),
               'shgan_fir_nhwc')


# ---- pointwise ---------------------------------------------------------------------------------------
@_on_tensor_device
def fromrgb(x, w, bias, wgain, act_alpha, act_gain, act_clamp, out):
    _f32c(x, 'x')
    n, ci, h, wd = x.shape
    lib = _lib.load()
    _lib.check(lib.shgan_fromrgb(_p(x), _p(w), _p(bias), wgain, act_alpha, act_gain, act_clamp, _p(out.hi), _p(out.lo),
                                n, ci, out.shape[3], h, wd, _stream()), 'shgan_fromrgb')
    return out


@_on_tensor_device
def fromrgb_masked(real, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp, out):
    """fromrgb over x = cat([mask - 0.5, real * mask]) formed on the fly (shgan_default.py:269-274); also writes x_out."""
    _f32c(real, 'real'); _f32c(mask, 'mask'); _f32c(x_out, 'x_out')
    n, c3, h, wd = real.shape
    lib = _lib.load()
    _lib.check(lib.shgan_fromrgb_masked(_p(real), _p(mask), _p(x_out), _p(w), _p(bias), wgain, act_alpha, act_gain, act_clamp,
                                       _p(out.hi), _p(out.lo), n, c3 + 1, out.shape[3], h, wd, _stream()), 'shgan_fromrgb_masked')
    return out


@_on_tensor_device
def torgb_combine(img_prev, rgb_partial, bias, f, img_out, comp_x=None, comp_out=None):
    n, _, h, w = img_out.shape
    lib = _lib.load()
    _lib.check(lib.shgan_torgb_combine(_p(img_prev), _p(rgb_partial), rgb_partial.shape[3], _p(bias), _p(f), _p(img_out),
                                      n, h, w, _p(comp_x), _p(comp_out), _stream()), 'shgan_torgb_combine')
    return img_out


@_on_tensor_device
def prepare_input(real, mask, out=None):
    """x = cat([mask - 0.5, real * mask]) (shgan_default.py:269-274).  real [N,3,H,W], mask [N,1,H,W] or [N,H,W]."""
    _f32c(real, 'real'); _f32c(mask, 'mask')
    n, _, h, w = real.shape
    if out is None:
        out = torch.empty((n, 4, h, w), dtype=torch.float32, device=real.device)
    lib = _lib.load()
    _lib.check(lib.shgan_prepare_input(_p(real), _p(mask), _p(out), n, h, w, _stream()), 'shgan_prepare_input')
    return out


@_on_tensor_device
def composite_cat(x, img, out=None):
    """cat([x[:, 0:1], x[:, 1:4]*m + img*(1-m)]), m = x[:, 0:1] + 0.5 (shgan_default.py:257-260): the discriminator input."""
    _f32c(x, 'x'); _f32c(img, 'img')
    n, _, h, w = x.shape
    if out is None:
        out = torch.empty((n, 4, h, w), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.shgan_composite_cat(_p(x), _p(img), _p(out), n, h, w, _stream()), 'shgan_composite_cat')
    return out


@_on_tensor_device
def mbstd_append(src, out, group_size):
    n, h, w, c = src.shape
    lib = _lib.load()
    _lib.check(lib.shgan_mbstd_append(_p(src.hi), _p(src.lo), _p(out.hi), _p(out.lo), n, h, w, c, out.shape[3], group_size,
                                     _stream()), 'shgan_mbstd_append')
    return out


# ---- dense / styles ------------------------------------------------------------------------------------
@_on_tensor_device
def dense(x0, w, bias, out, wgain, bgain=1.0, act=False, act_alpha=0.2, act_gain=SQRT2, act_clamp=256.0, x1=None):
    """out[b,o] = act((cat[x0,x1][b] . w[o]) * wgain + bias[o]*bgain).  x0/x1/out may be strided row views."""
    b, i0 = x0.shape
    i = i0 + (x1.shape[1] if x1 is not None else 0)
    o = w.shape[0]
    assert w.shape[1] == i and x0.stride(1) == 1 and out.stride(1) == 1
    lib = _lib.load()
    _lib.check(lib.shgan_dense_fwd(_p(x0), x0.stride(0), i0, _p(x1), x1.stride(0) if x1 is not None else 0, _p(w), _p(bias),
                                  _p(out), out.stride(0), b, i, o, wgain, bgain, 1 if act else 0, act_alpha, act_gain,
                                  act_clamp, _stream()), 'shgan_dense_fwd')
    return out


@_on_tensor_device
def normalize_2nd_moment(z, out=None):
    _f32c(z, 'z')
    out = torch.empty_like(z) if out is None else out
    lib = _lib.load()
    _lib.check(lib.shgan_normalize_2nd_moment(_p(z), _p(out), z.shape[0], z.shape[1], _stream()),
               'shgan_normalize_2nd_moment')
    return out


@_on_tensor_device
def style_prep(styles, wsq, s_hat, dcoef, demod, pre_scale=1.0):
    n, ci = styles.shape
    co = dcoef.shape[1] if dcoef is not None else 1
    lib = _lib.load()
    _lib.check(lib.shgan_style_prep(_p(styles), _p(wsq), _p(s_hat), _p(dcoef), n, ci, co, 1 if demod else 0, pre_scale,
                                   _stream()), 'shgan_style_prep')


@_on_tensor_device
def style_prep_batched(raw, layers):
    """raw fp32 [N, S] (row stride raw.stride(0)); layers: list of dicts(offset, ci, co, demod, pre_scale, wsq, s_hat, dcoef)."""
    tb = _lib.StyleBatch()
    tb.num_layers = len(layers)
    for i, L in enumerate(layers):
        tb.offset[i] = L['offset']; tb.ci[i] = L['ci']; tb.co[i] = L['co']; tb.demod[i] = 1 if L['demod'] else 0
        tb.pre_scale[i] = L['pre_scale']; tb.wsq[i] = _p(L['wsq']); tb.s_hat[i] = _p(L['s_hat']); tb.dcoef[i] = _p(L['dcoef'])
    lib = _lib.load()
    _lib.check(lib.shgan_style_prep_batched(_p(raw), raw.stride(0), raw.shape[0], C.byref(tb), _stream()), 'shgan_style_prep_batched')


# ---- SHU -----------------------------------------------------------------------------------------------
def shu_workspace_bytes(n, c, r):
    return int(_lib.load().shgan_shu_workspace_bytes(n, c, r))


@_on_tensor_device
def shu_pack(conv0_w, df1_w):
    """One-time fp16 hi/lo packing of the SHU weights for the tensor-core channel mix -> uint8 buffer (shgan_shu_pack)."""
    _f32c(conv0_w, 'conv0_w'); _f32c(df1_w, 'df1_w')
    c = conv0_w.shape[0] // 2
    lib = _lib.load()
    packed = torch.empty(int(lib.shgan_shu_packed_bytes(c)), dtype=torch.uint8, device=conv0_w.device)
    _lib.check(lib.shgan_shu_pack(_p(conv0_w), _p(df1_w), _p(packed), c, _stream()), 'shgan_shu_pack')
    return packed


@_on_tensor_device
def shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest_res, workspace=None, packed=None):
    """x fp32 [N,C,R,R]; outs: list of fp32 [N,C,r,r] for r = lowest_res*2^k; gauss: concatenated band masks;
    packed: shu_pack(conv0_w, df1_w) (None = pack on every call)."""
    _f32c(x, 'x')
    n, c, r, _ = x.shape
    if workspace is None:
        workspace = torch.empty(shu_workspace_bytes(n, c, r), dtype=torch.uint8, device=x.device)
    arr = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
    lib = _lib.load()
    _lib.check(lib.shgan_shu_fwd(_p(x), _p(conv0_w), _p(conv0_b), _p(df1_w), _p(cw), _p(gauss), _p(packed), _p(workspace), arr,
                                len(outs), n, c, r, lowest_res, _stream()), 'shgan_shu_fwd')
    return outs
