"""Import the UNMODIFIED reference model zoo from /root/reference (build container only).

The reference is Python, so it cannot travel to the GPU box; this helper exists only for
`make_golden.py` (fixture generation) and for the optional `test_oracle_vs_reference.py`
checks that are skipped when /root/reference is absent.

Two shims are needed: `matplotlib` (imported, never used on the path,
lib/model_zoo/common/utils.py:9) and `tensorboardX`-style optional deps of lib.log_service.
"""
import os
import sys
import types

REF_ROOT = os.environ.get('SHGAN_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'lib', 'model_zoo'))


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference `lib.model_zoo` package modules (stylegan, comodgan, shgan, upfirdn2d)."""
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REF_ROOT)
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
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import lib.model_zoo.stylegan as ref_stylegan
    import lib.model_zoo.comodgan as ref_comodgan
    import lib.model_zoo.shgan as ref_shgan
    from lib.model_zoo.stylegan_utils import upfirdn2d as ref_upfirdn2d
    from lib.model_zoo.stylegan_utils import conv2d_resample as ref_conv2d_resample
    from lib.model_zoo.common import utils as ref_utils
    return types.SimpleNamespace(
        stylegan=ref_stylegan, comodgan=ref_comodgan, shgan=ref_shgan,
        upfirdn2d=ref_upfirdn2d, conv2d_resample=ref_conv2d_resample, utils=ref_utils)
