"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/shgan_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'shgan_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(shgan_[a-z0-9_]+)\s*\(', hdr)))


def test_library_builds_loads_and_exports_header_symbols():
    from shgan_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    decl = _declared_symbols()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == decl                      # the ctypes binding covers exactly the header
    loaded = _lib.load()
    assert loaded.shgan_abi_version() == _lib.ABI_VERSION
    assert loaded.shgan_conv_num_nblocks(512, 0) == 16 and loaded.shgan_conv_num_nblocks(64, 0) == 2
    assert loaded.shgan_shu_workspace_bytes(2, 16, 64) == 2 * 2 * 32 * 64 * 33 * 4 + 256      # spec1 + spec2 (+ alignment slack)
    assert loaded.shgan_shu_packed_bytes(32) == 16384 + 3 * 32768                               # swizzled conv0 + three anchor-pair operand tiles, hi + lo
    assert loaded.shgan_shu_workspace_bytes(2, 32, 64) > 2 * 2 * 64 * 64 * 33 * 4             # + tensor-core mix operands (C == 32)
    # struct layouts the binding assumes
    assert ctypes.sizeof(_lib.Epilogue) % 8 == 0 and ctypes.sizeof(_lib.ConvDesc) % 8 == 0


def test_struct_layout_matches_c_compiler(tmp_path):
    """sizeof/offsetof of the two descriptor structs as seen by gcc == ctypes."""
    from shgan_b200 import _lib
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "shgan_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(shgan_epilogue), sizeof(shgan_conv_desc), offsetof(shgan_conv_desc, epi), offsetof(shgan_conv_desc, tap_w),'
                   'offsetof(shgan_epilogue, out_f32), offsetof(shgan_conv_desc, impl));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    exp = [ctypes.sizeof(_lib.Epilogue), ctypes.sizeof(_lib.ConvDesc), _lib.ConvDesc.epi.offset, _lib.ConvDesc.tap_w.offset,
           _lib.Epilogue.out_f32.offset, _lib.ConvDesc.impl.offset]
    assert got == exp


def test_sass_has_tcgen05_and_tma():
    """The convolution kernel really is a tcgen05/TMA kernel: SASS shows UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld)
    and UTMALDG (TMA tensor loads); see /opt/skills/guides/B200_PROFILING.md."""
    from shgan_b200 import build
    obj = os.path.join(os.path.dirname(build.build()), 'obj', 'conv_tc.o')
    cuobjdump = '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', obj], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG'):
        assert mnemonic in sass, mnemonic
    # issue paths run under elect.sync: no per-active-lane waterfall loop around the MMA / TMA instructions
    assert 'BRA.U.ANY' not in sass
    # the two-SM kernel: cta_group::2 MMAs, 2-CTA TMA loads, multicast commit, cluster barrier
    pair = subprocess.run([cuobjdump, '-sass', os.path.join(os.path.dirname(obj), 'conv_pair.o')], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA.2CTA', 'UTMALDG.4D.2CTA', 'UTCBAR.2CTA.MULTICAST', 'UCGABAR_ARV', 'STG.E.ENL2.256'):
        assert mnemonic in pair, mnemonic
    halo = subprocess.run([cuobjdump, '-sass', os.path.join(os.path.dirname(obj), 'conv_halo.o')], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in halo and 'BRA.U.ANY' not in halo
    # the SHU: tcgen05 channel mix with bulk-copied weight image, no generic shared-memory accesses, no scalar conversions;
    # register-FFT kernels moving whole planes with bulk copies, their shared-memory traffic as LDS / STS
    mix = subprocess.run([cuobjdump, '-sass', os.path.join(os.path.dirname(obj), 'shu_mix_tc.o')], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UBLKCP', 'F2FP'):
        assert mnemonic in mix, mnemonic
    assert 'LD.E.128' not in mix and 'ST.E.128' not in mix and 'F2F.F16.F32' not in mix and 'BRA.U.ANY' not in mix
    fft = subprocess.run([cuobjdump, '-sass', os.path.join(os.path.dirname(obj), 'shu_fft64.o')], capture_output=True, text=True).stdout
    assert 'UBLKCP' in fft and 'LDS' in fft and 'STS' in fft and 'ST.E.64' not in fft


def test_product_library_has_no_cuda_core_convolution():
    """The fp32 FMA cross-check kernel (impl=1) lives only in the test-only libshgan_b200_check.so."""
    import subprocess
    from shgan_b200 import _lib, build
    build.build()
    prod = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    chk = subprocess.run(['nm', '-D', '--defined-only', _lib.CHECK_LIB_PATH], capture_output=True, text=True).stdout
    assert 'conv_simt' not in prod and 'shgan_check_conv_igemm' not in prod
    assert 'shgan_check_conv_igemm' in chk and 'conv_simt' in chk
