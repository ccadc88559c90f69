"""SHU alone at one input resolution (development aid / ncu target):  python tools/shu_bench.py [R] [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P  # noqa: E402

dev = 'cuda'
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ch, lowest = 32, 4
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
ws = torch.empty(K.shu_workspace_bytes(n, ch, r), dtype=torch.uint8, device=dev)
packed = K.shu_pack(conv0_w, df1_w)
for _ in range(3):
    K.shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest, workspace=ws, packed=packed)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    K.shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest, workspace=ws, packed=packed)
e.record(); e.synchronize()
byts = n * ch * 4 * (r * r + sum(k * k for k in reslist))
ms = s.elapsed_time(e) / 5
print(f'R {r} N {n}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s')
