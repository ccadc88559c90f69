"""CPU emulation of the C-ABI kernels' DOCUMENTED semantics (include/shgan_b200.h), used only by the `not gpu`
tests to check the host-side logic of shgan_b200 (tap tables, weight packing, style/demod wiring, epilogue
wiring, ws indexing, SHU slicing) against the oracle and golden fixtures without a GPU.

It monkeypatches `shgan_b200.kernels` with torch-CPU functions that follow the header comments literally.
It is test infrastructure: the product never imports it, and the GPU tests never use it.
"""
import math

import numpy as np
import torch

from shgan_b200 import kernels as K
from shgan_b200.kernels import Planes


def _split(v):
    hi = v.to(torch.float16)
    lo = (v - hi.float()).to(torch.float16)
    return hi, lo


def make_epilogue(**kw):
    d = dict(dcoef=None, wgain=1.0, noise=None, noise_sn=0, noise_strength=None, bias=None, act=False, act_alpha=0.2,
             act_gain=1.0, act_clamp=-1.0, skip=None, next_scale=None, rgb_w=None, rgb_style=None, rgb_out=None, out=None,
             out_f32=None)
    d.update(kw)
    return d


def _apply_epi(acc, e, block_n, parity_split=False):
    n, oh, ow, co = acc.shape
    v = acc.clone()
    if e['dcoef'] is not None:
        v = v * e['dcoef'][:, None, None, :]
    v = v * e['wgain']
    if e['noise'] is not None:
        nz = e['noise'].reshape(-1)
        if e['noise_sn'] == 0:
            nzm = nz[:oh * ow].reshape(1, oh, ow, 1)
        else:
            assert e['noise_sn'] == oh * ow
            nzm = nz.reshape(n, oh, ow, 1)
        v = v + nzm * e['noise_strength'].reshape(())
    if e['bias'] is not None:
        v = v + e['bias']
    if e['act']:
        v = torch.where(v >= 0, v, v * e['act_alpha']) * e['act_gain']
        if e['act_clamp'] > 0:
            v = v.clamp(-e['act_clamp'], e['act_clamp'])
    else:
        v = v * e['act_gain']
    if e['skip'] is not None:
        v = v + e['skip'].float()
    if e['rgb_w'] is not None:
        t = v * e['rgb_style'][:, None, None, :]
        rb = 32                               # partial sums per block of 32 channels (include/shgan_b200.h)
        nblk = co // rb
        part = torch.einsum('nyxbc,jbc->nyxbj', t.reshape(n, oh, ow, nblk, rb), e['rgb_w'].reshape(3, nblk, rb))
        e['rgb_out'].zero_()
        e['rgb_out'][..., :3] = part
    if e['out_f32'] is not None:
        e['out_f32'].copy_(v)
    if e['out'] is not None:
        if e['next_scale'] is not None:
            v = v * e['next_scale'][:, None, None, :]
        hi, lo = _split(v)
        if parity_split == 2:
            e['out'].hi.copy_(hi[:, ::2, ::2])
            e['out'].lo.copy_(lo[:, ::2, ::2])
        elif parity_split:
            ph, pw = (oh + 1) // 2, (ow + 1) // 2
            oh_, ol_ = e['out'].hi.view(4, n, ph, pw, co), e['out'].lo.view(4, n, ph, pw, co)
            for py in range(2):
                for px in range(2):
                    sh, sl = hi[:, py::2, px::2], lo[:, py::2, px::2]
                    oh_[py * 2 + px][:, :sh.shape[1], :sh.shape[2]] = sh
                    ol_[py * 2 + px][:, :sl.shape[1], :sl.shape[2]] = sl
        else:
            e['out'].hi.copy_(hi)
            e['out'].lo.copy_(lo)


def _block_n(co, block_n=0):
    return block_n or (256 if co >= 256 else (128 if co >= 128 else 64))


def conv_num_nblocks(co, block_n=0):
    return co // 32


def conv_igemm(srcs, w_hi, w_lo, taps, oh, ow, epi=None, raw=None, block_n=0, passes=3, impl=0, acc_comp=None, splitk=None):
    n, _, _, c = srcs[0].shape
    w = w_hi.float() + w_lo.float()          # [T, Co, C]
    co = w.shape[1]
    assert c % 64 == 0 and co % 64 == 0 and len(taps) <= 16 and len(srcs) <= 4
    acc = torch.zeros((n, oh, ow, co), dtype=torch.float32)
    for (s, dy, dx, tw) in taps:
        x = srcs[s].float()
        hs, ws_ = x.shape[1], x.shape[2]
        # gather src[n, oy+dy, ox+dx] with zero fill
        win = torch.zeros((n, oh, ow, c), dtype=torch.float32)
        y0, y1 = max(0, -dy), min(oh, hs - dy)
        x0, x1 = max(0, -dx), min(ow, ws_ - dx)
        if y1 > y0 and x1 > x0:
            win[:, y0:y1, x0:x1] = x[:, y0 + dy:y1 + dy, x0 + dx:x1 + dx]
        acc += win @ w[tw].T
    if raw is not None:
        z, zsy, zsx, zoy, zox = raw
        z[:, zoy:zoy + (oh - 1) * zsy + 1:zsy, zox:zox + (ow - 1) * zsx + 1:zsx] = acc
    else:
        _apply_epi(acc, epi, _block_n(co, block_n))


def conv_up2(src, w_hi, w_lo, fx, fy, gain, epi, passes=3, acc_comp=None):
    """shgan_conv_up2 following the header text literally: z[n,2i+ky,2j+kx,o] += src[n,i,j,c] * w[ky*3+kx,o,c], then
    out[y,x] = gain * sum fy[a] fx[b] z[y+a-1, x+b-1] (zero outside), then the epilogue."""
    x = src.float()
    n, h, w, c = x.shape
    blocks = w_hi.shape[0]
    co = blocks * 64
    wp = w_hi.float() + w_lo.float()                      # [blk, slot, 64, C], slot -> tap {3,0,1,4,5,2,6,7,8}
    order = (3, 0, 1, 4, 5, 2, 6, 7, 8)
    wt = torch.zeros((9, co, c), dtype=torch.float32)
    for slot, tap in enumerate(order):
        wt[tap] = wp[:, slot].reshape(co, c)
    z = torch.zeros((n, 2 * h + 1 + 2, 2 * w + 1 + 2, co), dtype=torch.float32)      # one zero ring = z outside its support
    for ky in range(3):
        for kx in range(3):
            z[:, 1 + ky:1 + ky + 2 * h:2, 1 + kx:1 + kx + 2 * w:2] += x @ wt[ky * 3 + kx].T
    acc = torch.zeros((n, 2 * h, 2 * w, co), dtype=torch.float32)
    for a in range(4):
        for b in range(4):
            acc += float(fy[a]) * float(fx[b]) * gain * z[:, a:a + 2 * h, b:b + 2 * w]
    _apply_epi(acc, epi, 64)


def fir_nhwc(src, f, gain, pads, epi, parity_split=False, rank1=False):
    x = src.float() if isinstance(src, Planes) else src
    n, ih, iw, c = x.shape
    px0, px1, py0, py1 = pads
    assert tuple(f.shape) == (4, 4) and min(pads) >= 0
    xp = torch.nn.functional.pad(x, (0, 0, px0, px1, py0, py1))
    oh, ow = ih + py0 + py1 - 3, iw + px0 + px1 - 3
    acc = torch.zeros((n, oh, ow, c), dtype=torch.float32)
    for i in range(4):
        for j in range(4):
            acc += float(f[i, j]) * gain * xp[:, i:i + oh, j:j + ow]
    _apply_epi(acc, epi, c, parity_split=parity_split)


def nchw_to_planes(x, add=None, scale=None, out=None, c_off=0):
    n, c, h, w = x.shape
    if out is None:
        out = Planes.empty(n, h, w, c, x.device)
    v = x.permute(0, 2, 3, 1).float()
    if add is not None:
        v = v + add.float()[..., c_off:c_off + c]
    if scale is not None:
        v = v * scale[:, None, None, :]
    hi, lo = _split(v)
    out.hi[..., c_off:c_off + c] = hi
    out.lo[..., c_off:c_off + c] = lo
    return out


def planes_to_nchw(p, c_off=0, c=None, out=None):
    c_tot = p.shape[3]
    c = c_tot - c_off if c is None else c
    v = p.float()[..., c_off:c_off + c].permute(0, 3, 1, 2).contiguous()
    if out is not None:
        out.copy_(v)
        return out
    return v


def planes_add_nchw(p, x, c_off):
    return nchw_to_planes(x, add=p, out=p, c_off=c_off)


def planes_add_nchw_multi(planes_list, xs, c_off_list):
    for p, x, co in zip(planes_list, xs, c_off_list):
        planes_add_nchw(p, x, co)


def nhwc_to_nchw_f32(x):
    return x.permute(0, 3, 1, 2).contiguous()


def fromrgb(x, w, bias, wgain, act_alpha, act_gain, act_clamp, out):
    v = torch.einsum('nchw,oc->nhwo', x, w * wgain) + bias
    v = torch.where(v >= 0, v, v * act_alpha) * act_gain
    if act_clamp > 0:
        v = v.clamp(-act_clamp, act_clamp)
    hi, lo = _split(v)
    out.hi.copy_(hi)
    out.lo.copy_(lo)
    return out


def fromrgb_masked(real, mask, x_out, w, bias, wgain, act_alpha, act_gain, act_clamp, out):
    x_out.copy_(torch.cat([mask - 0.5, real * mask], dim=1))
    return fromrgb(x_out, w, bias, wgain, act_alpha, act_gain, act_clamp, out)


def torgb_combine(img_prev, rgb_partial, bias, f, img_out, comp_x=None, comp_out=None):
    n, _, h, w = img_out.shape
    v = rgb_partial[..., :3].sum(3).permute(0, 3, 1, 2) + bias.view(1, 3, 1, 1)
    if img_prev is not None:
        from oracle import shgan_oracle as O
        up = O.upsample2d(img_prev.numpy(), f.numpy())
        v = torch.from_numpy(up) + v
    img_out.copy_(v)
    if comp_x is not None:
        m = comp_x[:, 0:1] + 0.5
        o = comp_x[:, 1:4] * m + img_out * (1 - m)
        comp_out.copy_((o * 127.5 + 127.5).clamp(0, 255).to(torch.uint8))
    return img_out


def mbstd_append(src, out, group_size):
    from oracle import shgan_oracle as O
    x = src.float().permute(0, 3, 1, 2).numpy()
    y = torch.from_numpy(O.minibatch_std(x, group_size, 1)).permute(0, 2, 3, 1)
    full = torch.zeros(out.shape, dtype=torch.float32)
    full[..., :y.shape[3]] = y
    hi, lo = _split(full)
    out.hi.copy_(hi)
    out.lo.copy_(lo)
    return out


def dense(x0, w, bias, out, wgain, bgain=1.0, act=False, act_alpha=0.2, act_gain=math.sqrt(2.0), act_clamp=256.0, x1=None):
    x = x0 if x1 is None else torch.cat([x0, x1], dim=1)
    v = (x @ w.T) * wgain
    if bias is not None:
        v = v + bias * bgain
    if act:
        v = torch.where(v >= 0, v, v * act_alpha) * act_gain
        if act_clamp > 0:
            v = v.clamp(-act_clamp, act_clamp)
    out.copy_(v)
    return out


def normalize_2nd_moment(z, out=None):
    v = z * (z.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    if out is not None:
        out.copy_(v)
        return out
    return v


def style_prep(styles, wsq, s_hat, dcoef, demod, pre_scale=1.0):
    if demod:
        s = styles * styles.square().mean().rsqrt()
        s_hat.copy_(s)
        dcoef.copy_((s.square() @ wsq.T + 1e-8).rsqrt())
    else:
        s_hat.copy_(styles * pre_scale)


def style_prep_batched(raw, layers):
    for L in layers:
        style_prep(raw[:, L['offset']:L['offset'] + L['ci']], L['wsq'], L['s_hat'], L['dcoef'], L['demod'], L['pre_scale'])


def shu_workspace_bytes(n, c, r):
    return 16


def shu_pack(conv0_w, df1_w):
    return torch.zeros(16, dtype=torch.uint8)


def shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest_res, workspace=None, packed=None):
    """Emulated with the oracle's SHU; checks that the constants handed to the kernel are the oracle's."""
    from oracle import shgan_oracle as O
    n, c, r, _ = x.shape
    sd = {'s.conv0.weight': conv0_w.numpy().reshape(2 * c, 2 * c, 1, 1), 's.conv0.bias': conv0_b.numpy(), 's.df1.weight': df1_w.numpy()}
    np.testing.assert_allclose(cw.numpy(), O.make_cweight((2, 3), (r, r // 2 + 1)), atol=1e-7)
    gm = O.gaussian_weight_maps(r, lowest_res)
    np.testing.assert_allclose(gauss.numpy(), np.concatenate([gm[k].reshape(-1) for k in sorted(gm)]), atol=1e-7)
    res = O.shu_forward(sd, x.numpy(), prefix='s', input_res=r, lowest_res=lowest_res)
    for o, k in zip(outs, sorted(res)):
        o.copy_(torch.from_numpy(res[k]))
    return outs


def install(monkeypatch):
    """Patch shgan_b200.kernels (and the engine's device check) with the CPU emulation."""
    import shgan_b200.engine as E
    for name in ['make_epilogue', 'conv_num_nblocks', 'conv_igemm', 'conv_up2', 'fir_nhwc', 'nchw_to_planes', 'planes_to_nchw',
                 'planes_add_nchw', 'planes_add_nchw_multi', 'nhwc_to_nchw_f32', 'fromrgb', 'fromrgb_masked', 'torgb_combine', 'mbstd_append', 'dense', 'normalize_2nd_moment',
                 'style_prep', 'style_prep_batched', 'shu_workspace_bytes', 'shu_pack', 'shu_fwd']:
        monkeypatch.setattr(K, name, globals()[name])
    monkeypatch.setattr(E, '_check_device', lambda dev: None)
