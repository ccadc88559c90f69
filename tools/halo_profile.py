"""Development aid: per-role cycle accounting of conv_halo_kernel (library built with SHGAN_NVCC_FLAGS=-DSHGAN_HALO_PROFILE).
Runs single plain 3x3 layers of the 512^2 generator at batch 16 and prints, averaged over the CTAs, the cycles the MMA
issuer spent blocked on (a_full, w_full, t_empty) and the epilogue warps on (t_full wait, chunk drain, final epilogue)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P, _lib  # noqa: E402

LAYERS = [(64, 64, 512), (128, 128, 256), (512, 512, 64)]
lib = _lib.load()
for ci, co, r in LAYERS:
    n = 16
    x = torch.randn(n, r, r, ci, device='cuda')
    xp = K.Planes.empty(n, r, r, ci, x.device)
    hi = x.half(); xp.hi.copy_(hi); xp.lo.copy_((x - hi.float()).half())
    w = torch.randn(co, ci, 3, 3, device='cuda')
    wh, wl = P.pack_conv_weight(w)
    out = K.Planes.empty(n, r, r, co, x.device)
    bias = torch.randn(co, device='cuda')
    dco = torch.rand(n, co, device='cuda') + 0.5
    for impl in (3,):
        epi = K.make_epilogue(dcoef=dco, bias=bias, act=True, act_alpha=0.2, act_gain=1.414, act_clamp=256.0, next_scale=dco, out=out)
        for _ in range(2):
            K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), r, r, epi=epi, impl=impl)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), r, r, epi=epi, impl=impl)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        buf = (ctypes.c_longlong * (148 * 16))()
        rc = lib.shgan_debug_halo_profile(buf)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
        m = a.mean(axis=0)
        tf = 2 * n * r * r * ci * co * 9 / ms / 1e9
        print(f'C{ci}->{co} @{r}: {ms:.3f} ms  {tf:.0f} TFLOP/s alg | MMA issuer: total {m[0]:.0f} clk, wait a_full {m[1]/m[0]:.1%} w_full {m[2]/m[0]:.1%} '
              f't_empty {m[3]/m[0]:.1%} | epilogue(warp4 / warp11): total {m[4]:.0f}, t_full wait {m[5]/m[4]:.1%} / {m[9]/m[8]:.1%}, drain {m[6]/m[4]:.1%} / {m[10]/m[8]:.1%}, '
              f'final {m[7]/m[4]:.1%} / {m[11]/m[8]:.1%}')
