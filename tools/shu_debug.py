"""Stage-by-stage check of the input_res-64 SHU path on the GPU (development aid): forward spectrum, channel mix and band
outputs, each against float64 numpy computed from the reference value of the previous stage.  python tools/shu_debug.py [N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P  # noqa: E402

dev = 'cuda'
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
r, ch, lowest = 64, 32, 4
bins = r * (r // 2 + 1)
masks = P.gaussian_band_masks(r, lowest, 3, False)
reslist = sorted(masks)
g = torch.Generator().manual_seed(5)
conv0_w = (torch.randn(64, 64, generator=g) / 8)
conv0_b = (torch.randn(64, generator=g) * 0.1)
df1_w = (1 / 64 + 0.1 / 64 * torch.randn(64, 384, generator=g))
cw = P.make_cweight((2, 3), (r, r // 2 + 1)).contiguous()
gauss = torch.cat([masks[k].reshape(-1) for k in reslist]).contiguous()
x = torch.randn(n, ch, r, r, generator=g)
outs = [torch.empty(n, ch, k, k, device=dev) for k in reslist]
ws = torch.zeros(K.shu_workspace_bytes(n, ch, r), dtype=torch.uint8, device=dev)
packed = K.shu_pack(conv0_w.to(dev), df1_w.to(dev))
K.shu_fwd(x.to(dev), conv0_w.to(dev), conv0_b.to(dev), df1_w.to(dev), cw.to(dev), gauss.to(dev), outs, lowest, workspace=ws, packed=packed)
torch.cuda.synchronize()
spec = ws[:2 * n * 64 * bins * 4].view(torch.float32).reshape(2, n, 64, r // 2 + 1, r).cpu().numpy().astype(np.float64)   # [kx][s]


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


xs = x.numpy().astype(np.float64)
f = np.fft.rfft2(xs, norm='forward')
f = np.concatenate([f[:, :, r // 2 + 1:], f[:, :, :r // 2 + 1]], axis=2)            # [n, c, s, kx]
s1 = np.concatenate([f.real, f.imag], axis=1)                                      # [n, 64, s, kx]
print('spec1 rel err', rel(spec[0].transpose(0, 1, 3, 2), s1))
w0, b0, d1 = conv0_w.numpy().astype(np.float64), conv0_b.numpy().astype(np.float64), df1_w.numpy().astype(np.float64)
t = np.maximum(np.einsum('oi,nisk->nosk', w0, s1) + b0[None, :, None, None], 0)
y = np.einsum('io,nisk->nosk', d1, t).reshape(n, 64, 6, r, r // 2 + 1)
s2 = (y * cw.numpy().astype(np.float64)[None, None]).sum(2)
print('spec2 rel err', rel(spec[1].transpose(0, 1, 3, 2), s2))
if rel(spec[1].transpose(0, 1, 3, 2), s2) > 1e-4:
    d = np.abs(spec[1].transpose(0, 1, 3, 2) - s2)
    print('  worst per sample', d.max(axis=(1, 2, 3)), 'per channel block', d.reshape(n, 2, 32, r, -1).max(axis=(0, 2, 3, 4)))
    print('  worst per kx', d.max(axis=(0, 1, 2)).round(6))
fc = s2[:, :32] + 1j * s2[:, 32:]
for bi, k in enumerate(reslist):
    sp = fc[:, :, r // 2 - k // 2:r // 2 + k // 2, :k // 2 + 1] * masks[k].numpy().astype(np.float64)[None, None]
    sp = np.concatenate([sp[:, :, k - k // 2 - 1:], sp[:, :, :k - k // 2 - 1]], axis=2)
    ref = np.fft.irfft2(sp, s=(k, k), norm='forward')
    print(f'band {k} rel err', rel(outs[bi].cpu().numpy().astype(np.float64), ref))
