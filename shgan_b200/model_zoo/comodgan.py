"""Co-modulated GAN generator with the reference's module API (lib/model_zoo/comodgan.py:30-481).

Blocks are parameter containers with the reference's names/shapes; `Encoder.forward`, `Synthesis.forward` and
`Generator.forward` execute through `shgan_b200.engine.GeneratorEngine` (fused sm_100a kernel sequence) and
convert to the reference's NCHW fp32 tensors only at their own boundary.
"""
import copy
import weakref

import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from ..engine import GeneratorEngine
from .common.get_model import get_model, register
from .stylegan import Discriminator as Discriminator_StyleGan
from .stylegan import Generator as Generator_StyleGan
from .stylegan import Mapping as Mapping_StyleGan
from .stylegan import conv2d_layer, dense, synthesis_layer, torgb_layer
from .. import packing as P

version = '0'
symbol = 'comodgan'

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


@register('comodgan_mapping')
class Mapping(Mapping_StyleGan):
    pass


class encoder_block(nn.Module):
    """fromrgb (first block) -> conv0 3x3 (= skip feature) -> conv1 blur + 3x3 stride 2
    (stylegan.py:624-656 discrim_block as used by comodgan.py:34-64; reslink is not used by SH-GAN)."""

    def __init__(self, ic_n, mc_n, oc_n, rgb_n=None, resample_filter=[1, 3, 3, 1], activation=ACT, reslink=False,
                 use_fp16=False):
        super().__init__()
        if reslink or use_fp16:
            raise NotImplementedError('reslink / fp16 encoder blocks are not used by the released SH-GAN configs')
        self.register_buffer('resample_filter', P.setup_filter(resample_filter))
        self.fromrgb = None
        if rgb_n is not None:
            self.fromrgb = conv2d_layer(rgb_n, mc_n, 1, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.conv0 = conv2d_layer(ic_n, mc_n, 3, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.conv1 = conv2d_layer(mc_n, oc_n, 3, bias=True, activation=activation, up=1, down=2, resample_filter=resample_filter)
        self.reslink = reslink
        self.use_fp16 = use_fp16

    def forward(self, x, img):
        if self.fromrgb is not None:
            y = self.fromrgb(img.float())
            x = x + y if x is not None else y
        feat = self.conv0(x)
        return self.conv1(feat), None, feat


class encoder_epilogue(nn.Module):
    """4x4 block: conv 3x3 (= skip feature) -> flatten -> dense -> dropout (comodgan.py:66-113)."""

    def __init__(self, ic_n, oc_n, resolution, cmap_dim, rgb_n=None, mbstd_group_size=4, mbstd_c_n=1, activation=ACT,
                 reslink=True, use_dropout=True, has_extra_final_layer=True):
        super().__init__()
        if rgb_n is not None or mbstd_c_n > 0 or cmap_dim is not None:
            raise NotImplementedError('fromrgb / minibatch-std / cmap in the encoder epilogue are not used by SH-GAN')
        self.ic_n, self.cmap_dim, self.resolution, self.rgb_n, self.reslink = ic_n, cmap_dim, resolution, rgb_n, reslink
        self.fromrgb = None
        self.mbstd = None
        self.conv = conv2d_layer(ic_n, ic_n, 3, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.fc = dense(ic_n * (resolution ** 2), oc_n, activation=activation)
        self.out = dense(oc_n, oc_n, activation=None) if has_extra_final_layer else None
        self.dropout = nn.Dropout(p=0.5) if use_dropout else None

    def forward(self, x, img=None, cmap=None):
        feat = self.conv(x.float())
        x = self.fc(feat.flatten(1))
        if self.out is not None:
            x = self.out(x)
        if self.dropout is not None:
            x = self.dropout(x)
        return x, feat


class _EngineMixin:
    """Gives a sub-module access to the fused engine of the generator that owns it (or a private one)."""

    def _engine(self):
        owner = self.__dict__.get('_owner')
        G = owner() if owner is not None else None
        if G is None or (getattr(G, 'encoder', None) is not self and getattr(G, 'synthesis', None) is not self):
            raise RuntimeError('this module runs through a comodgan_generator; construct Generator(mapping, encoder, '
                               'synthesis) and call it (or its .encoder/.synthesis) instead')
        return G.engine()

    def __deepcopy__(self, memo):
        # the back-reference to the owning generator must not be copied (a weakref deep-copies atomically and would keep
        # pointing at the ORIGINAL generator's engine and parameters): a copied sub-module is unattached until a
        # Generator adopts it (Generator.__init__ / Generator.__deepcopy__)
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != '_owner':
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new


@register('comodgan_encoder', version)
class Encoder(nn.Module, _EngineMixin):
    """Conv encoder producing the global code and the per-resolution skip features (comodgan.py:115-205)."""

    def __init__(self, resolution=256, ic_n=3, oc_n=1024, ch_base=16384, ch_max=512, use_fp16_before_res=16,
                 resample_filter=[1, 3, 3, 1], activation=ACT, mbstd_group_size=4, mbstd_c_n=1, c_dim=None, cmap_dim=None,
                 use_dropout=True, has_extra_final_layer=True):
        super().__init__()
        log2res = int(np.log2(resolution))
        if 2 ** log2res != resolution:
            raise ValueError
        if use_fp16_before_res is not None:
            raise NotImplementedError('the released SH-GAN configs run the encoder in fp32 (use_fp16_before_res: null)')
        if c_dim is not None and c_dim > 0:
            raise NotImplementedError('conditional encoder')
        self.encode_res = [2 ** i for i in range(log2res, 1, -1)]
        self.ic_n, self.ch_base, self.ch_max = ic_n, ch_base, ch_max
        self.resample_filter, self.activation = resample_filter, activation
        for idx, (ri, rj) in enumerate(zip(self.encode_res[:-1], self.encode_res[1:])):
            ci, cj = min(ch_base // ri, ch_max), min(ch_base // rj, ch_max)
            setattr(self, f'b{ri}', encoder_block(ci, ci, cj, rgb_n=ic_n if idx == 0 else None, resample_filter=resample_filter,
                                                  activation=activation, reslink=False, use_fp16=False))
        self.mapping = None
        hidden = min(ch_base // self.encode_res[-1], ch_max)
        self.b4 = encoder_epilogue(hidden, oc_n, resolution=4, cmap_dim=None, activation=activation,
                                   mbstd_group_size=mbstd_group_size, mbstd_c_n=mbstd_c_n, reslink=False,
                                   use_dropout=use_dropout, has_extra_final_layer=has_extra_final_layer)

    def forward(self, img, c=None):
        """-> (x_global [N,oc_n], feats {res: NCHW fp32}) like the reference (comodgan.py:190-205)."""
        if self.training and self.b4.dropout is not None:
            raise NotImplementedError('training-mode dropout: shgan_b200 implements the eval forward only')
        x_global, feats = self._engine().encoder(img)
        return x_global.clone(), {r: K.planes_to_nchw(p) for r, p in feats.items()}


class synthesis_block_first(nn.Module):
    """4x4 block: dense(x_global) + feats[4] -> modulated conv -> torgb (comodgan.py:207-262)."""

    def __init__(self, w0_dim, oc_n, w_dim, resolution, rgb_n=None, activation=ACT):
        super().__init__()
        self.resolution = resolution
        self.fc = dense(w0_dim, oc_n * (resolution ** 2), activation=activation)
        self.num_conv, self.num_torgb = 1, 0
        self.conv = synthesis_layer(oc_n, oc_n, 3, w0_dim + w_dim, resolution=4, bias=True, activation=activation)
        self.torgb = None
        if rgb_n is not None:
            self.torgb = torgb_layer(oc_n, rgb_n, 1, w0_dim + w_dim, activation=None)
            self.num_torgb += 1


class synthesis_block(nn.Module):
    """conv0 (up 2) + feats[res] -> conv1 -> img = upsample(img) + torgb (comodgan.py:264-340)."""

    def __init__(self, ic_n, oc_n, w_dim, w0_dim, resolution, rgb_n, resample_filter=[1, 3, 3, 1], activation=ACT,
                 res_link=False, use_fp16=False):
        super().__init__()
        if ic_n == 0:
            raise ValueError
        if res_link or use_fp16:
            raise NotImplementedError('res_link / fp16 synthesis blocks are not used by the released SH-GAN configs')
        self.w_dim, self.resolution, self.use_fp16, self.res_link = w_dim, resolution, use_fp16, res_link
        self.register_buffer('resample_filter', P.setup_filter(resample_filter))
        self.num_conv, self.num_torgb = 2, 0
        self.const = None
        self.conv0 = synthesis_layer(ic_n, oc_n, 3, w_dim=w_dim + w0_dim, resolution=resolution, up=2, activation=activation,
                                     resample_filter=resample_filter, use_noise=True)
        self.conv1 = synthesis_layer(oc_n, oc_n, 3, w_dim=w_dim + w0_dim, resolution=resolution, up=1, activation=activation,
                                     resample_filter=None, use_noise=True)
        self.torgb = None
        if rgb_n is not None:
            self.torgb = torgb_layer(oc_n, rgb_n, 1, w_dim=w_dim + w0_dim, activation=None)
            self.num_torgb += 1


@register('comodgan_synthesis', version)
class Synthesis(nn.Module, _EngineMixin):
    """Co-modulated synthesis network (comodgan.py:342-433)."""

    def __init__(self, w_dim=512, w0_dim=1024, resolution=256, rgb_n=3, ch_base=16384, ch_max=512, use_fp16_after_res=16,
                 resample_filter=[1, 3, 3, 1], activation=ACT):
        super().__init__()
        log2res = int(np.log2(resolution))
        if 2 ** log2res != resolution:
            raise ValueError
        if use_fp16_after_res is not None:
            raise NotImplementedError('the released SH-GAN configs run the synthesis in fp32 (use_fp16_after_res: null)')
        if rgb_n is None or rgb_n > 3:
            raise NotImplementedError('rgb_n must be 1..3')
        self.w_dim, self.resolution, self.rgb_n = w_dim, resolution, rgb_n
        self.block_res = [2 ** i for i in range(2, log2res + 1)]
        self.activation = activation
        # 14 / 16 / 18 for 256 / 512 / 1024 as hard-coded by the reference (comodgan.py:362-367); same rule elsewhere
        self.num_ws = 2 * log2res - 2
        hidden = min(ch_base // self.block_res[0], ch_max)
        self.b4 = synthesis_block_first(w0_dim, hidden, w_dim, resolution=4, rgb_n=rgb_n, activation=activation)
        for ri, rj in zip(self.block_res[:-1], self.block_res[1:]):
            ci, cj = min(ch_base // ri, ch_max), min(ch_base // rj, ch_max)
            setattr(self, f'b{rj}', synthesis_block(ci, cj, w_dim=w_dim, w0_dim=w0_dim, resolution=rj, rgb_n=rgb_n,
                                                    resample_filter=resample_filter, activation=activation))

    def forward(self, x, feats, ws, noise_mode='random'):
        """x = x_global [N,w0_dim], feats {res: NCHW fp32}, ws [N,num_ws,w_dim] -> img [N,rgb_n,R,R]."""
        eng = self._engine()
        eng._ensure()
        pl = {r: K.nchw_to_planes(f.contiguous().float(), out=eng._planes(f'ext.feat{r}', f.shape[0], f.shape[2], f.shape[3], f.shape[1]))
              for r, f in feats.items()}
        img = eng.synthesis(x.contiguous().float(), pl, ws.to(torch.float32).contiguous(), noise_mode=noise_mode)
        return img[:, :self.rgb_n].clone()


@register('comodgan_generator', version)
class Generator(Generator_StyleGan):
    """mapping + encoder + synthesis (comodgan.py:435-481).  forward(x, z, c, noise_mode) is one fused pass."""

    def __init__(self, mapping, encoder, synthesis):
        super().__init__(mapping, synthesis)
        self.encoder = encoder if isinstance(encoder, nn.Module) else get_model()(encoder)
        self.ic_n = self.encoder.ic_n
        self.__dict__['_engine_obj'] = None
        self._adopt()

    def _adopt(self):
        # sub-modules called on their own (G.encoder(x), G.synthesis(...)) run through this generator's engine
        for sub in (self.encoder, self.synthesis):
            sub.__dict__['_owner'] = weakref.ref(self)

    def engine(self, passes=None, impl=None, graphs=None):
        eng = self.__dict__.get('_engine_obj')
        if eng is None:
            eng = GeneratorEngine(self)
            self.__dict__['_engine_obj'] = eng
        if passes is not None:
            eng.passes = passes
        if impl is not None:
            eng.impl = impl
        if graphs is not None:
            eng.graphs = graphs
        return eng

    def __deepcopy__(self, memo):
        # the engine holds device buffers keyed on this instance's parameters: a copy builds its own lazily
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_engine_obj' else copy.deepcopy(v, memo)
        new._adopt()
        return new

    def forward(self, x, z, c=None, truncation_psi=1, truncation_cutoff=None, noise_mode='random'):
        """x: [N,4,R,R] (mask-0.5, masked rgb); z: [N,z_dim]; c: [N,0] (unused) -> img [N,3,R,R] fp32."""
        if truncation_psi != 1:
            raise NotImplementedError('truncation is not used by the SH-GAN eval path (truncation_psi=1)')
        if self.training and self.encoder.b4.dropout is not None:
            raise NotImplementedError('training-mode dropout: shgan_b200 implements the eval forward only')
        img = self.engine().forward(x, z, noise_mode=noise_mode)
        return img[:, :self.img_channels].clone()

    def forward_inpaint(self, real, mask, z, noise_mode='random'):
        """The eval loop's whole per-batch device work in one fused pass (lib/experiments/shgan_default.py:257-274):
        x = cat([mask - 0.5, real * mask]) is formed inside the fromrgb kernel, the composite + uint8 quantisation inside
        the last torgb kernel.  real [N,3,R,R] in [-1,1], mask [N,1,R,R] in {0,1} -> (img fp32, composite uint8)."""
        img, comp = self.engine().forward((real, mask), z, noise_mode=noise_mode, composite=True)
        return img.clone(), comp

    def forward_composite(self, x, z, noise_mode='random'):
        """Generator forward fused with the eval loop's composite + uint8 quantisation
        (lib/experiments/shgan_default.py:257-262) -> (img fp32, composite uint8)."""
        img, comp = self.engine().forward(x, z, noise_mode=noise_mode, composite=True)
        return img.clone(), comp


@register('comodgan_discriminator', version)
class Discriminator(Discriminator_StyleGan):
    """comodgan.py:483-485."""
    pass
