"""StyleGAN2 building blocks of the generator path with the reference's module API
(lib/model_zoo/stylegan.py:28-430, 582-606): identical constructor arguments, attribute names, parameter /
buffer names and shapes -- so a reference `state_dict` loads with strict=True -- while every forward runs on
the sm_100a kernels of this package (`shgan_b200.ops` at operator level; the whole-generator call is fused
by `shgan_b200.engine.GeneratorEngine`).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import ops
from .. import packing as P
from .common.get_model import get_model, register

version = '0'
symbol = 'stylegan'

modulated_conv2d = ops.modulated_conv2d


class conv2d(nn.Conv2d):
    """nn.Conv2d with StyleGAN weight scaling (stylegan.py:28-64).  Only used by the SHU's 1x1 conv0, which the
    fused SHU kernel consumes as a [out,in] matrix; the op-level forward handles 1x1/stride 1."""

    def __init__(self, *args, **kwargs):
        use_wscale = kwargs.pop('use_wscale', False)
        super().__init__(*args, **kwargs)
        in_channels = args[0] if len(args) > 0 else kwargs['in_channels']
        kernel_size = args[2] if len(args) > 2 else kwargs['kernel_size']
        he_std = 1 / np.sqrt(in_channels * kernel_size * kernel_size)
        self.weight_gain = he_std if use_wscale else 1
        self.bias_gain = 1
        nn.init.normal_(self.weight, mean=0.0, std=1 if use_wscale else he_std)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    def forward(self, x):
        if self.kernel_size != (1, 1) or self.stride != (1, 1) or self.padding != (0, 0) or self.groups != 1:
            raise NotImplementedError('only the 1x1 stride-1 conv2d of the SHU is on the generator path')
        y = ops.conv2d_resample(x, (self.weight * self.weight_gain).to(x.dtype), padding=0,
                                _epi=dict(bias=self.bias.detach().float()) if self.bias is not None else None)
        return y


class dense(nn.Module):
    """Fully connected layer (stylegan.py:66-101)."""

    def __init__(self, in_features, out_features, bias=True, bias_init=0, activation=None, lr_multi=1):
        super().__init__()
        self.activation_spec = activation
        self.activation = P.parse_activation(activation)
        self.weight = nn.Parameter(torch.randn([out_features, in_features]) / lr_multi)
        self.bias = nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multi / np.sqrt(in_features)
        self.bias_gain = lr_multi
        self.repr = 'dense({}, {}, bias={}, act={}, lr_multi={})'.format(in_features, out_features, bias, activation, lr_multi)

    def forward(self, x):
        x = x.contiguous().float()
        out = torch.empty((x.shape[0], self.weight.shape[0]), dtype=torch.float32, device=x.device)
        a = self.activation
        K.dense(x, self.weight.detach(), None if self.bias is None else self.bias.detach(), out, float(self.weight_gain),
                float(self.bias_gain), a is not None, a['alpha'] if a else 0.0, a['gain'] if a else 1.0,
                (a['clamp'] if a and a['clamp'] is not None else -1.0))
        return out

    def __repr__(self):
        return self.repr


class conv2d_layer(nn.Module):
    """Conv layer with optional up/down-sampling, bias and lrelu_agc (stylegan.py:195-241)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation=None, up=1, down=1,
                 resample_filter=[1, 3, 3, 1]):
        super().__init__()
        self.up = up
        self.down = down
        if resample_filter is not None:
            self.register_buffer('resample_filter', P.setup_filter(resample_filter))
        else:
            self.resample_filter = None
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.activation_spec = activation
        self.activation = P.parse_activation(activation)
        self.weight = nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = nn.Parameter(torch.zeros([out_channels])) if bias else None
        self.repr = 'conv2d_layer({}, {}, kernal_size={}, bias={}, up={}, down={}, act={})'.format(
            in_channels, out_channels, kernel_size, bias, up, down, activation)

    def _act_epi(self, gain):
        a = self.activation
        epi = dict(bias=None if self.bias is None else self.bias.detach().float())
        if a is not None:
            epi.update(act=True, act_alpha=a['alpha'], act_gain=a['gain'] * gain,
                       act_clamp=(a['clamp'] * gain if a['clamp'] is not None else -1.0))
        else:
            epi.update(act=False, act_gain=float(gain))
        return epi

    def forward(self, x, gain=1):
        w = (self.weight * self.weight_gain).to(x.dtype)
        return ops.conv2d_resample(x, w, f=self.resample_filter, up=self.up, down=self.down, padding=self.padding,
                                   flip_weight=(self.up == 1), _epi=self._act_epi(gain))

    def __repr__(self):
        return self.repr


class synthesis_layer(conv2d_layer):
    """Modulated conv + noise + bias + lrelu_agc (stylegan.py:243-304)."""

    def __init__(self, in_channels, out_channels, kernel_size, w_dim, resolution, bias=True,
                 activation='lrelu_agc(alpha=0.2, gain=sqrt_2)', up=1, resample_filter=[1, 3, 3, 1], use_noise=True):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, activation=activation, up=up, down=1,
                         resample_filter=resample_filter)
        self.affine = dense(w_dim, in_channels, bias=True, bias_init=1, activation=None)
        self.resolution = resolution
        self.use_noise = use_noise
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, fused_modconv=True, gain=1, noise_mode='random'):
        assert noise_mode in ['random', 'const', 'none']
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        y = modulated_conv2d(x=x, weight=self.weight.detach(), styles=styles, noise=noise, up=self.up, padding=self.padding,
                             resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
        # bias + activation on an already materialised NCHW tensor (operator-level call; the fused engine does this in-kernel)
        y = y + self.bias.detach().view(1, -1, 1, 1)
        a = self.activation
        if a is not None:
            y = ops.lrelu_agc(y, a['alpha'], a['gain'], a['clamp'], extra_gain=gain)
        else:
            y = y * gain
        return y


class torgb_layer(conv2d_layer):
    """1x1 modulated conv without demodulation (stylegan.py:306-337)."""

    def __init__(self, in_channels, out_channels, kernel_size, w_dim, activation=None):
        super().__init__(in_channels, out_channels, kernel_size, bias=True, activation=activation, up=1, down=1,
                         resample_filter=None)
        self.affine = dense(w_dim, in_channels, bias=True, bias_init=1, activation=None)

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        y = modulated_conv2d(x=x, weight=self.weight.detach(), styles=styles, demodulate=False, fused_modconv=fused_modconv)
        y = y + self.bias.detach().view(1, -1, 1, 1)
        if self.activation is not None:
            a = self.activation
            y = ops.lrelu_agc(y, a['alpha'], a['gain'], a['clamp'])
        return y


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    assert dim == 1 and eps == 1e-8 and x.ndim == 2
    return K.normalize_2nd_moment(x.contiguous().float())


@register('stylegan2_mapping', version)
class Mapping(nn.Module):
    """z -> w mapping network (stylegan.py:346-430); conditioning (c_dim > 0) is not on the SH-GAN path."""

    def __init__(self, z_dim=512, c_dim=0, w_dim=512, num_ws=14, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)', lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        if c_dim != 0:
            raise NotImplementedError('class-conditional mapping (c_dim > 0) is not used by SH-GAN')
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers, self.w_avg_beta = num_layers, w_avg_beta
        if layer_features is None:
            layer_features = w_dim
        features = [z_dim] + [layer_features] * (num_layers - 1) + [w_dim]
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', dense(features[idx], features[idx + 1], activation=activation, lr_multi=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        x = normalize_2nd_moment(z.to(torch.float32))
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class discrim_block(nn.Module):
    """(fromrgb) -> conv0 -> conv1 (down 2) with an optional 1x1 down-sampling skip branch (stylegan.py:624-684).
    Parameter container; the arithmetic runs in engine.GeneratorEngine / engine.DiscriminatorEngine."""

    def __init__(self, ic_n, mc_n, oc_n, rgb_n=None, resample_filter=[1, 3, 3, 1],
                 activation='lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)', reslink=False, use_fp16=False):
        super().__init__()
        if use_fp16:
            raise NotImplementedError('fp16 blocks are not used by the released SH-GAN configs')
        self.register_buffer('resample_filter', P.setup_filter(resample_filter))
        self.fromrgb = None
        if rgb_n is not None:
            self.fromrgb = conv2d_layer(rgb_n, mc_n, 1, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.conv0 = conv2d_layer(ic_n, mc_n, 3, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.conv1 = conv2d_layer(mc_n, oc_n, 3, bias=True, activation=activation, up=1, down=2, resample_filter=resample_filter)
        self.reslink = reslink
        if reslink:
            self.skip = conv2d_layer(mc_n, oc_n, 1, bias=False, activation=None, up=1, down=2, resample_filter=resample_filter)
        self.use_fp16 = use_fp16


class minibatch_std_layer(nn.Module):
    """stylegan.py:686-705 (parameter-free; runs as shgan_mbstd_append)."""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels


class discrim_epilogue(nn.Module):
    """mbstd -> conv 3x3 -> fc -> out (stylegan.py:707-755)."""

    def __init__(self, ic_n, resolution, cmap_dim, rgb_n=None, mbstd_group_size=4, mbstd_c_n=1,
                 activation='lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)', reslink=True):
        super().__init__()
        if rgb_n is not None or cmap_dim is not None:
            raise NotImplementedError('fromrgb / cmap in the discriminator epilogue are not used by SH-GAN')
        self.ic_n, self.cmap_dim, self.resolution, self.rgb_n, self.reslink = ic_n, cmap_dim, resolution, rgb_n, reslink
        self.fromrgb = None
        self.mbstd = minibatch_std_layer(group_size=mbstd_group_size, num_channels=mbstd_c_n) if mbstd_c_n > 0 else None
        self.conv = conv2d_layer(ic_n + mbstd_c_n, ic_n, 3, bias=True, activation=activation, up=1, down=1, resample_filter=None)
        self.fc = dense(ic_n * (resolution ** 2), ic_n, activation=activation)
        self.out = dense(ic_n, 1, activation=None)


@register('stylegan2_discriminator', version)
class Discriminator(nn.Module):
    """StyleGAN2 residual discriminator (stylegan.py:757-838); forward(img, c) -> logits [N,1]."""

    def __init__(self, resolution=256, ic_n=3, ch_base=16384, ch_max=512, use_fp16_before_res=16, resample_filter=[1, 3, 3, 1],
                 activation='lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)', mbstd_group_size=4, mbstd_c_n=1, c_dim=None,
                 cmap_dim=None):
        super().__init__()
        log2res = int(np.log2(resolution))
        if 2 ** log2res != resolution:
            raise ValueError
        if use_fp16_before_res is not None:
            raise NotImplementedError('the released SH-GAN configs run the discriminator in fp32 (use_fp16_before_res: null)')
        if c_dim is not None and c_dim > 0:
            raise NotImplementedError('conditional discriminator')
        self.encode_res = [2 ** i for i in range(log2res, 1, -1)]
        self.ic_n, self.ch_base, self.ch_max = ic_n, ch_base, ch_max
        self.resample_filter, self.activation = resample_filter, activation
        for idx, (ri, rj) in enumerate(zip(self.encode_res[:-1], self.encode_res[1:])):
            ci, cj = min(ch_base // ri, ch_max), min(ch_base // rj, ch_max)
            setattr(self, f'b{ri}', discrim_block(ci, ci, cj, rgb_n=ic_n if idx == 0 else None, resample_filter=resample_filter,
                                                  activation=activation, reslink=True, use_fp16=False))
        self.mapping = None
        self.b4 = discrim_epilogue(min(ch_base // self.encode_res[-1], ch_max), resolution=4, cmap_dim=None, activation=activation,
                                   mbstd_group_size=mbstd_group_size, mbstd_c_n=mbstd_c_n)
        self.__dict__['_engine_obj'] = None

    def engine(self, passes=None, impl=None):
        from ..engine import DiscriminatorEngine
        eng = self.__dict__.get('_engine_obj')
        if eng is None:
            eng = DiscriminatorEngine(self)
            self.__dict__['_engine_obj'] = eng
        if passes is not None:
            eng.passes = passes
        if impl is not None:
            eng.impl = impl
        return eng

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_engine_obj' else copy.deepcopy(v, memo)
        return new

    def forward(self, img, c=None, **kwargs):
        return self.engine().forward(img)


@register('stylegan2_generator', version)
class Generator(nn.Module):
    """mapping + synthesis container (stylegan.py:582-606); sub-configs or ready modules are accepted."""

    def __init__(self, mapping, synthesis):
        super().__init__()
        self.mapping = mapping if isinstance(mapping, nn.Module) else get_model()(mapping)
        self.synthesis = synthesis if isinstance(synthesis, nn.Module) else get_model()(synthesis)
        if self.synthesis.num_ws != self.mapping.num_ws:
            raise ValueError
        self.num_ws = self.mapping.num_ws
        self.z_dim = self.mapping.z_dim
        self.c_dim = self.mapping.c_dim
        self.w_dim = self.mapping.w_dim
        self.img_resolution = self.synthesis.resolution
        self.img_channels = self.synthesis.rgb_n
