"""ctypes binding of the C-ABI library declared in include/shgan_b200.h.

The product path has NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is
raised with the library's own error message (`shgan_last_error`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libshgan_b200.so')

SHGAN_MAX_TAPS = 16
SHGAN_MAX_SRC = 4
ABI_VERSION = 3

vp = C.c_void_p
fp = C.c_void_p  # float* passed as raw device addresses
i32 = C.c_int
i64 = C.c_int64
f32 = C.c_float


class Epilogue(C.Structure):
    """shgan_epilogue (include/shgan_b200.h)."""
    _fields_ = [
        ('dcoef', fp), ('wgain', f32), ('noise', fp), ('noise_sn', i64), ('noise_strength', fp), ('bias', fp),
        ('act', i32), ('act_alpha', f32), ('act_gain', f32), ('act_clamp', f32),
        ('skip_hi', vp), ('skip_lo', vp), ('next_scale', fp), ('rgb_w', fp), ('rgb_style', fp), ('rgb_out', fp),
        ('out_hi', vp), ('out_lo', vp), ('out_f32', fp),
    ]


class ConvDesc(C.Structure):
    """shgan_conv_desc (include/shgan_b200.h)."""
    _fields_ = [
        ('num_src', i32), ('src_hi', vp * SHGAN_MAX_SRC), ('src_lo', vp * SHGAN_MAX_SRC),
        ('src_h', i32 * SHGAN_MAX_SRC), ('src_w', i32 * SHGAN_MAX_SRC),
        ('N', i32), ('C', i32), ('Co', i32), ('w_hi', vp), ('w_lo', vp), ('w_taps', i32), ('ntaps', i32),
        ('tap_src', i32 * SHGAN_MAX_TAPS), ('tap_dy', i32 * SHGAN_MAX_TAPS), ('tap_dx', i32 * SHGAN_MAX_TAPS),
        ('tap_w', i32 * SHGAN_MAX_TAPS),
        ('OH', i32), ('OW', i32), ('mode', i32), ('z', fp), ('ZH', i32), ('ZW', i32), ('zsy', i32), ('zsx', i32),
        ('zoy', i32), ('zox', i32), ('epi', Epilogue), ('block_n', i32), ('passes', i32), ('impl', i32), ('acc_comp', f32),
    ]


class Up2Desc(C.Structure):
    """shgan_up2_desc (include/shgan_b200.h)."""
    _fields_ = [
        ('src_hi', vp), ('src_lo', vp), ('N', i32), ('H', i32), ('W', i32), ('C', i32), ('Co', i32),
        ('w_hi', vp), ('w_lo', vp), ('fx', f32 * 4), ('fy', f32 * 4), ('gain', f32), ('epi', Epilogue),
        ('passes', i32), ('acc_comp', f32),
    ]


SHGAN_MAX_ADD = 8


class AddBatch(C.Structure):
    """shgan_add_batch (include/shgan_b200.h)."""
    _fields_ = [
        ('num', i32), ('hi', vp * SHGAN_MAX_ADD), ('lo', vp * SHGAN_MAX_ADD), ('x', vp * SHGAN_MAX_ADD),
        ('hw', i32 * SHGAN_MAX_ADD), ('c_off', i32 * SHGAN_MAX_ADD), ('c_tot', i32 * SHGAN_MAX_ADD),
        ('work_start', C.c_longlong * (SHGAN_MAX_ADD + 1)),
    ]


SHGAN_MAX_STYLE_LAYERS = 40


class StyleBatch(C.Structure):
    """shgan_style_batch (include/shgan_b200.h)."""
    _fields_ = [
        ('num_layers', i32), ('offset', i64 * SHGAN_MAX_STYLE_LAYERS),
        ('ci', i32 * SHGAN_MAX_STYLE_LAYERS), ('co', i32 * SHGAN_MAX_STYLE_LAYERS), ('demod', i32 * SHGAN_MAX_STYLE_LAYERS),
        ('pre_scale', f32 * SHGAN_MAX_STYLE_LAYERS), ('wsq', vp * SHGAN_MAX_STYLE_LAYERS),
        ('s_hat', vp * SHGAN_MAX_STYLE_LAYERS), ('dcoef', vp * SHGAN_MAX_STYLE_LAYERS),
        ('block_start', i32 * (SHGAN_MAX_STYLE_LAYERS + 1)),
    ]


# name -> (restype, argtypes); every symbol include/shgan_b200.h declares
SIGNATURES = {
    'shgan_abi_version': (i32, []),
    'shgan_last_error': (C.c_char_p, []),
    'shgan_launch_count': (C.c_uint64, []),
    'shgan_upfirdn2d_fwd': (i32, [fp, fp, fp] + [i32] * 14 + [i32, f32, vp]),
    'shgan_nchw_to_planes': (i32, [fp, vp, vp, fp, vp, vp] + [i32] * 6 + [vp]),
    'shgan_planes_to_nchw': (i32, [vp, vp, fp] + [i32] * 6 + [vp]),
    'shgan_planes_add_nchw': (i32, [vp, vp, fp] + [i32] * 6 + [vp]),
    'shgan_planes_add_nchw_multi': (i32, [C.POINTER(AddBatch), i32, i32, vp]),
    'shgan_nhwc_to_nchw_f32': (i32, [fp, fp] + [i32] * 4 + [vp]),
    'shgan_conv_igemm': (i32, [C.POINTER(ConvDesc), vp]),
    'shgan_conv_num_nblocks': (i32, [i32, i32]),
    'shgan_conv_up2': (i32, [C.POINTER(Up2Desc), vp]),
    'shgan_fir_nhwc': (i32, [fp, vp, vp, fp, i32, i32, f32] + [i32] * 8 + [C.POINTER(Epilogue), i32, vp]),
    'shgan_fromrgb': (i32, [fp, fp, fp, f32, f32, f32, f32, vp, vp] + [i32] * 5 + [vp]),
    'shgan_fromrgb_masked': (i32, [fp, fp, fp, fp, fp, f32, f32, f32, f32, vp, vp] + [i32] * 5 + [vp]),
    'shgan_torgb_combine': (i32, [fp, fp, i32, fp, fp, fp, i32, i32, i32, fp, vp, vp]),
    'shgan_prepare_input': (i32, [fp, fp, fp, i32, i32, i32, vp]),
    'shgan_composite_cat': (i32, [fp, fp, fp, i32, i32, i32, vp]),
    'shgan_mbstd_append': (i32, [vp, vp, vp, vp] + [i32] * 6 + [vp]),
    'shgan_dense_fwd': (i32, [fp, i64, i32, fp, i64, fp, fp, fp, i64, i32, i32, i32, f32, f32, i32, f32, f32, f32, vp]),
    'shgan_normalize_2nd_moment': (i32, [fp, fp, i32, i32, vp]),
    'shgan_style_prep': (i32, [fp, fp, fp, fp, i32, i32, i32, i32, f32, vp]),
    'shgan_style_prep_batched': (i32, [fp, i64, i32, C.POINTER(StyleBatch), vp]),
    'shgan_shu_packed_bytes': (i64, [i32]),
    'shgan_shu_pack': (i32, [fp, fp, vp, i32, vp]),
    'shgan_shu_workspace_bytes': (i64, [i32, i32, i32]),
    'shgan_shu_fwd': (i32, [fp, fp, fp, fp, fp, fp, vp, vp, C.POINTER(fp), i32, i32, i32, i32, i32, vp]),
}

_lib = None


def load():
    """Load the library (once).  Raises if it has not been built: there is no CPU/PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m shgan_b200.build` (shgan_b200 has no fallback path)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.shgan_abi_version() != ABI_VERSION:
        raise RuntimeError(f'libshgan_b200.so ABI {lib.shgan_abi_version()} != binding ABI {ABI_VERSION}')
    _lib = lib
    return lib


_check_lib = None
CHECK_LIB_PATH = os.path.join(_HERE, 'lib', 'libshgan_b200_check.so')


def load_check():
    """TEST-ONLY library with the fp32 FMA cross-check convolution (`impl=1`); never loaded by the product path."""
    global _check_lib
    if _check_lib is None:
        if not os.path.exists(CHECK_LIB_PATH):
            raise RuntimeError(f'{CHECK_LIB_PATH} is missing: impl=1 is a test facility, build it with `python -m shgan_b200.build`')
        lib = C.CDLL(CHECK_LIB_PATH)
        lib.shgan_check_conv_igemm.restype = i32
        lib.shgan_check_conv_igemm.argtypes = [C.POINTER(ConvDesc), vp]
        lib.shgan_last_error.restype = C.c_char_p
        _check_lib = lib
    return _check_lib


def check(rc, what):
    if rc != 0:
        msg = load().shgan_last_error()
        raise RuntimeError(f'{what} failed (code {rc}): {msg.decode() if msg else "?"}')


def launch_count():
    return int(load().shgan_launch_count())
