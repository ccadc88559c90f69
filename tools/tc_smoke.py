"""Fast hang/garbage detector for the tcgen05 conv kernel: run under `timeout 60` before anything expensive."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import shgan_oracle as O
from shgan_b200 import kernels as K, packing as P
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
g = np.random.default_rng(0)
ok = True
for (n, ci, co, h, w) in [(1, 64, 64, 8, 16), (2, 128, 128, 16, 16), (1, 512, 256, 16, 16), (2, 64, 512, 33, 20)]:
    x = g.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = g.standard_normal((co, ci, 3, 3)).astype(np.float32)
    xp = K.nchw_to_planes(t(x)); wh, wl = P.pack_conv_weight(t(wt))
    y = torch.empty((n, h, w, co), device='cuda')
    K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), h, w, epi=K.make_epilogue(out_f32=y), passes=3, impl=0)
    torch.cuda.synchronize()
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), padding=1)
    e = np.abs(K.nhwc_to_nchw_f32(y).cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f'tc_smoke C{ci}->{co} {h}x{w}: rel err {e:.2e}', flush=True)
    ok &= e < 1e-5
sys.exit(0 if ok else 1)
