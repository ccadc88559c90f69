"""Cycle accounting of the SHU channel mix (SHGAN_MIX_TRACE=1): per-tile hand-off timestamps of CTA 0.  python tools/mix_trace.py"""
import os
import sys

import numpy as np
import torch

os.environ['SHGAN_MIX_TRACE'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P  # noqa: E402

dev = 'cuda'
r, n, ch, lowest = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 512, 32, 4
bins = r * (r // 2 + 1)
masks = P.gaussian_band_masks(r, lowest, 3, False)
reslist = sorted(masks)
g = torch.Generator().manual_seed(r)
conv0_w = (torch.randn(64, 64, generator=g) / 8).to(dev)
conv0_b = (torch.randn(64, generator=g) * 0.1).to(dev)
df1_w = (1 / 64 + 0.1 / 64 * torch.randn(64, 384, generator=g)).to(dev)
cw = P.make_cweight((2, 3), (r, r // 2 + 1)).to(dev).contiguous()
gauss = torch.cat([masks[k].reshape(-1) for k in reslist]).to(dev).contiguous()
x = torch.randn(n, ch, r, r, device=dev)
outs = [torch.empty(n, ch, k, k, device=dev) for k in reslist]
ws = torch.zeros(K.shu_workspace_bytes(n, ch, r), dtype=torch.uint8, device=dev)
packed = K.shu_pack(conv0_w, df1_w)
for _ in range(3):
    K.shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest, workspace=ws, packed=packed)
torch.cuda.synchronize()
base = ws.data_ptr()
a256 = lambda v: (v + 255) // 256 * 256
extra = a256(base + 2 * n * 64 * bins * 4)
cw_kx = a256(extra + 16384 + 3 * 32768)
off = cw_kx + 6 * bins * 4 - base
tr = ws[off:off + 4 * 32 * 8 * 8].view(torch.int64).reshape(4, 32, 8).cpu().numpy()
t0 = tr[0, 0, 0]
names = ['MMA: g1wait g1go tfullwait tfullgo g2a_go g2b_go tile_end', 'P: top xempty written | sts_done fence_done', 'E2: a_wait a_full a_done b_wait b_full b_done stores_done', 'E1: top d1full done | ld0 ld1 sts_done fence_done']
for role in range(4):
    print(names[role])
    for it in range(8, 14):
        print('  it', it, ' '.join('%7d' % (v - t0) if v else '      -' for v in tr[role, it]))
per = (tr[0, 28, 1] - tr[0, 8, 1]) / 20
print('tile period (cycles)', per)
