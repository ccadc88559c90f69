"""Import the UNMODIFIED reference model zoo.

Search order: $SHGAN_REFERENCE_ROOT, /root/reference (the build container), <repo>/baseline/_ref (a verbatim,
git-ignored copy made by baseline/install_reference.py -- the only one of the three that exists on the GPU box).
Used by `make_golden.py` (fixture generation), by the tests that compare against the live reference
(tests/test_gpu_reference.py, skipped when no tree is found) and by `bench.py --impl reference`.

Two import shims are needed: `matplotlib` (imported, never used on the path,
lib/model_zoo/common/utils.py:9) and `easydict` / `tensorboardX`-style optional deps of lib.cfg_helper / lib.log_service.
The reference's own sources are not touched.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


def _candidates():
    env = os.environ.get('SHGAN_REFERENCE_ROOT')
    return ([env] if env else []) + ['/root/reference', os.path.join(_REPO, 'baseline', '_ref')]


def reference_root():
    for c in _candidates():
        if os.path.isdir(os.path.join(c, 'lib', 'model_zoo')):
            return c
    return None


REF_ROOT = reference_root()


def reference_available():
    return reference_root() is not None


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference `lib.model_zoo` package modules (stylegan, comodgan, shgan, upfirdn2d, ...)."""
    root = reference_root()
    if root is None:
        raise RuntimeError('reference tree not found in any of ' + ', '.join(_candidates()))
    for name in ['matplotlib', 'matplotlib.pyplot', 'tensorboardX', 'easydict']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = _stub(name)
                if name == 'easydict':
                    class EasyDict(dict):
                        __getattr__ = dict.__getitem__
                        __setattr__ = dict.__setitem__
                    m.EasyDict = EasyDict
    if root not in sys.path:
        sys.path.insert(0, root)
    # the plugin JIT (custom_ops.py:46-124) builds under TORCH_EXTENSIONS_DIR; keep it inside the repo copy so that a box
    # without a writable home still works, and so that a prebuilt plugin travels with the snapshot
    os.environ.setdefault('TORCH_EXTENSIONS_DIR', os.path.join(_REPO, 'baseline', '_ref', 'torch_extensions'))
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    # custom_ops.get_plugin builds with torch.utils.cpp_extension.load and then calls importlib.import_module(name)
    # (custom_ops.py:111); torch >= 2 no longer leaves the build directory on sys.path, so put it there
    plug_dir = os.path.join(os.environ['TORCH_EXTENSIONS_DIR'], 'upfirdn2d_plugin')
    if plug_dir not in sys.path:
        sys.path.append(plug_dir)
    import lib.model_zoo.stylegan as ref_stylegan
    import lib.model_zoo.comodgan as ref_comodgan
    import lib.model_zoo.shgan as ref_shgan
    from lib.model_zoo.stylegan_utils import upfirdn2d as ref_upfirdn2d
    from lib.model_zoo.stylegan_utils import conv2d_resample as ref_conv2d_resample
    from lib.model_zoo.common import utils as ref_utils
    return types.SimpleNamespace(
        stylegan=ref_stylegan, comodgan=ref_comodgan, shgan=ref_shgan,
        upfirdn2d=ref_upfirdn2d, conv2d_resample=ref_conv2d_resample, utils=ref_utils, root=root)


def build_reference_generator(R, resolution, ch_base=32768, ch_max=512, num_ws=None):
    """`comodgan_generator` with the args of configs/model/{stylegan,comodgan,shgan}.yaml (shgan_g256 / shgan_g512),
    built through the reference's own constructors."""
    import math
    log2 = int(math.log2(resolution))
    if num_ws is None:
        num_ws = {256: 14, 512: 16, 1024: 18}.get(resolution, 2 * log2 - 2)
    m = R.comodgan.Mapping(z_dim=512, c_dim=0, w_dim=512, num_ws=num_ws, num_layers=8, embed_features=None,
                           layer_features=None, activation=ACT, lr_multiplier=0.01, w_avg_beta=0.995)
    e = R.shgan.Encoder(resolution=resolution, ic_n=4, oc_n=1024, ch_base=ch_base, ch_max=ch_max, use_fp16_before_res=None,
                        resample_filter=[1, 3, 3, 1], activation=ACT, mbstd_group_size=0, mbstd_c_n=0, c_dim=None,
                        cmap_dim=None, use_dropout=True, has_extra_final_layer=False, shu_channels=32, shu_df_freedom=[2, 3],
                        shu_df_type='piecewise_linear', shu_input_res=64, shu_lowest_res=4, shu_tail_sigma_mult=3,
                        shu_gaussian_at_input_res=False)
    s = R.comodgan.Synthesis(w_dim=512, w0_dim=1024, resolution=resolution, rgb_n=3, ch_base=ch_base, ch_max=ch_max,
                             use_fp16_after_res=None, resample_filter=[1, 3, 3, 1], activation=ACT)
    if not hasattr(s, 'num_ws'):
        s.num_ws = num_ws  # comodgan.py:362-367 only defines it for 256/512/1024
    return R.comodgan.Generator(m, e, s).eval().requires_grad_(False)


def build_reference_discriminator(R, resolution, ch_base=32768, ch_max=512):
    return R.comodgan.Discriminator(resolution=resolution, ic_n=4, ch_base=ch_base, ch_max=ch_max,
                                    use_fp16_before_res=None).eval().requires_grad_(False)
