"""Development aid: per-role cycle accounting of conv_pair_kernel (library built with SHGAN_NVCC_FLAGS=-DSHGAN_PAIR_PROFILE)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P, _lib  # noqa: E402
lib = _lib.load()
for ci, co, r in [(256, 256, 128), (512, 512, 64), (512, 512, 32)]:
    n = 16
    x = torch.randn(n, r, r, ci, device='cuda')
    xp = K.Planes.empty(n, r, r, ci, x.device)
    hi = x.half(); xp.hi.copy_(hi); xp.lo.copy_((x - hi.float()).half())
    wh, wl = P.pack_conv_weight(torch.randn(co, ci, 3, 3, device='cuda'))
    out = K.Planes.empty(n, r, r, co, x.device)
    bias = torch.randn(co, device='cuda'); dco = torch.rand(n, co, device='cuda') + 0.5
    epi = K.make_epilogue(dcoef=dco, bias=bias, act=True, act_alpha=0.2, act_gain=1.414, act_clamp=256.0, next_scale=dco, out=out)
    for _ in range(3):
        K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), r, r, epi=epi, impl=4)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), r, r, epi=epi, impl=4); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    buf = (ctypes.c_longlong * (148 * 16))()
    lib.shgan_debug_pair_profile(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
    lead, peer = a[0::2].mean(axis=0), a[1::2].mean(axis=0)
    print(f'C{ci}->{co} @{r}: {ms:.3f} ms {2*n*r*r*ci*co*9/ms/1e9:.0f} TF/s | issuer total {lead[0]:.0f} clk: wait full {lead[1]/lead[0]:.1%} tempty {lead[2]/lead[0]:.1%} | '
          f'epilogue leader: total {lead[4]:.0f}, tfull wait {lead[5]/lead[4]:.1%} drain {lead[6]/lead[4]:.1%} final {lead[7]/lead[4]:.1%} | '
          f'peer: tfull wait {peer[5]/peer[4]:.1%} drain {peer[6]/peer[4]:.1%} final {peer[7]/peer[4]:.1%}')
