"""What pure streams reach on this GPU (development aid): memset / fill / copy / reduction over 1 GiB,
CUDA events.  Measured on B200: fill 7.45 TB/s, reduction 5.9 TB/s, copy 6.5 TB/s (read + write), byte memset 3.9 TB/s."""
import torch
d='cuda'
a=torch.empty(1<<30, dtype=torch.uint8, device=d); b=torch.empty(1<<30, dtype=torch.uint8, device=d)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e)/n
ms=t(lambda: a.zero_()); print(f'memset 1 GiB: {ms:.3f} ms  {a.numel()/ms/1e6:.0f} GB/s written')
ms=t(lambda: b.copy_(a)); print(f'copy 1 GiB: {ms:.3f} ms  {2*a.numel()/ms/1e6:.0f} GB/s read+written')
af=a.view(torch.float32)
ms=t(lambda: af.fill_(1.5)); print(f'fill f32 1 GiB: {ms:.3f} ms  {a.numel()/ms/1e6:.0f} GB/s written')
ms=t(lambda: af.sum()); print(f'sum (read) 1 GiB: {ms:.3f} ms  {a.numel()/ms/1e6:.0f} GB/s read')
