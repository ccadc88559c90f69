"""SH-GAN encoder and Spectral Hint Unit with the reference's module API (lib/model_zoo/shgan.py:70-383).

`SHU.forward(x) -> {res: tensor}` runs the cuFFT-free fused kernels (shgan_shu_fwd): rFFT2 -> shifted
(re|im) 1x1 conv + ReLU -> heterogeneous per-band filter -> Gaussian band split -> per-band irFFT2.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import packing as P
from .comodgan import Encoder as Encoder_base
from .common.get_model import register
from .stylegan import conv2d

version = '0'
symbol = 'shgan'


def make_cweight(half_size, half_sample, type='piecewise_linear', oddeven_aligned=True, device='cpu'):
    """Blend weights [fh*fw, H, W/2+1] of the heterogeneous filter (shgan.py:70-121)."""
    if type != 'piecewise_linear' or not oddeven_aligned:
        raise NotImplementedError("only type='piecewise_linear', oddeven_aligned=True is used by SH-GAN")
    return P.make_cweight(list(half_size), list(half_sample)).to(device)


class heterogeneous_filter(nn.Module):
    """Per-frequency-bin linear map whose matrix is a piecewise-linear blend of fh*fw anchor matrices
    (shgan.py:123-160).  Holds the parameter; the arithmetic lives in the fused SHU kernel."""

    def __init__(self, in_channels, out_channels, freedom, type, init='ones'):
        super().__init__()
        if type != 'piecewise_linear':
            raise NotImplementedError("dfilter_type must be 'piecewise_linear'")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.freedom, self.type = list(freedom), type
        fh, fw = self.freedom
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels * fh * fw))
        if init == 'ones':
            nn.init.ones_(self.weight)


class SHU(nn.Module):
    """Spectral Hint Unit (shgan.py:252-336)."""

    def __init__(self, in_channels, out_channels, dfilter_freedom=[3, 2], dfilter_type='piecewise_linear', input_res=256,
                 lowest_res=4, tail_sigma_mult=3, gaussian_at_input_res=False):
        super().__init__()
        if in_channels != out_channels:
            raise NotImplementedError('the SHU adds its output back into its input channels: in == out')
        if list(dfilter_freedom) != [2, 3]:
            raise NotImplementedError('the fused SHU kernel is specialised for dfilter_freedom=[2, 3] (shgan.yaml)')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.input_res, self.lowest_res = input_res, lowest_res
        self.conv0 = conv2d(in_channels * 2, in_channels * 2, 1, 1, 0)
        self.df1 = heterogeneous_filter(in_channels * 2, out_channels * 2, freedom=dfilter_freedom, type=dfilter_type)
        nn.init.normal_(self.df1.weight, mean=1 / (out_channels * 2), std=0.1 / (out_channels * 2))
        self.tail_sigma_mult = tail_sigma_mult
        self.gaussian_at_input_res = gaussian_at_input_res
        self.reslist = [2 ** i for i in range(int(np.log2(lowest_res)), int(np.log2(input_res)) + 1)]
        # plain dict of CPU tensors like the reference (not part of the state_dict)
        self.gaussian_weight_map = P.gaussian_band_masks(input_res, lowest_res, tail_sigma_mult, gaussian_at_input_res)
        self.__dict__['_consts'] = None

    def _device_consts(self, device):
        c = self.__dict__.get('_consts')
        if c is None or c[0] != device:
            r = self.input_res
            cw = P.make_cweight(self.df1.freedom, (r, r // 2 + 1)).to(device).contiguous()
            gauss = torch.cat([self.gaussian_weight_map[k].reshape(-1) for k in self.reslist]).to(device).contiguous()
            c = (device, cw, gauss)
            self.__dict__['_consts'] = c
        return c[1], c[2]

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('SHU.forward: shgan_b200 has no CPU path; expected a CUDA tensor')
        n, c, h, w = x.shape
        assert c == self.in_channels and h == self.input_res and w == self.input_res
        cw, gauss = self._device_consts(x.device)
        c2 = 2 * c
        outs = [torch.empty((n, c, r, r), dtype=torch.float32, device=x.device) for r in self.reslist]
        K.shu_fwd(x.contiguous().float(), self.conv0.weight.detach().reshape(c2, c2).contiguous(), self.conv0.bias.detach(),
                  self.df1.weight.detach().contiguous(), cw, gauss, outs, self.lowest_res)
        return {r: o for r, o in zip(self.reslist, outs)}


@register('shgan_encoder', version)
class Encoder(Encoder_base):
    """CoModGAN encoder + SHU on the last `shu_channels` channels of the `shu_input_res` feature (shgan.py:338-383)."""

    def __init__(self, *args, **kwargs):
        self.shu_input_res = kwargs.pop('shu_input_res')
        self.shu_lowest_res = kwargs.pop('shu_lowest_res')
        self.shu_channels = kwargs.pop('shu_channels')
        self.shu_df_freedom = kwargs.pop('shu_df_freedom')
        self.shu_df_type = kwargs.pop('shu_df_type')
        self.shu_tail_sigma_mult = kwargs.pop('shu_tail_sigma_mult')
        self.shu_gaussian_at_input_res = kwargs.pop('shu_gaussian_at_input_res')
        super().__init__(*args, **kwargs)
        self.shu = SHU(self.shu_channels, self.shu_channels, self.shu_df_freedom, self.shu_df_type,
                       input_res=self.shu_input_res, lowest_res=self.shu_lowest_res,
                       tail_sigma_mult=self.shu_tail_sigma_mult, gaussian_at_input_res=self.shu_gaussian_at_input_res)
