"""Device timings of the BASELINE.json configs that bench.py's default run does not cover (bench.py = C3, FFHQ-512 batch 16):

  C2  FFHQ-256 generator forward, batch 32, 1 GPU
  C4  Places2-512 generator + discriminator step (forward only: the reference ships no training step), batch 8 per GPU:
      G(x, z) -> composite -> D(cat[mask - 0.5, fake])
  C5  SHU-only sweep over the input resolution (the kernel covers 4..128; the released model uses 64): algorithmic GB/s
      against the HBM copy peak plus the channel-mix GFLOP/s

One JSON line per config on stdout (same keys as bench.py where they apply).  CUDA-event timing, >= 3 warm-ups; every
working set except the small SHU sizes is far larger than L2, the SHU loop flushes L2 between iterations.

    python tools/bench_configs.py [c2] [c4] [c5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shgan_b200 import kernels as K, packing as P, synthetic as S  # noqa: E402
from shgan_b200.model_zoo import get_model  # noqa: E402

GFLOP_G = {512: 238.785, 256: 180.635}
GFLOP_D512 = 123.204
HBM_FALLBACK = 6650.0


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return HBM_FALLBACK, 'fallback'


def time_steps(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def c2():
    res, batch = 256, 32
    G = S.random_generator(res, seed=0)
    x, z = S.synthetic_batch(batch, res, seed=1)
    x, z = x.cuda(), z.cuda()
    ms = time_steps(lambda: G.forward_composite(x, z, noise_mode='random'))
    print(json.dumps(dict(metric='256x256 inpaint images/sec', value=batch / ms * 1e3, unit='images/s', n_gpus=1, ms_per_step=ms,
                          config=dict(workload='C2: FFHQ-256 shgan_ffhq256_eval generator forward, batch 32, synthetic images/masks, random-init weights',
                                      noise_mode='random', cuda_graph=True),
                          whole_step_algorithmic_tflops=GFLOP_G[res] * batch / ms)))


def c4():
    res, batch = 512, 8
    G = S.random_generator(res, seed=0)
    torch.manual_seed(3)
    D = get_model()(dict(type='comodgan_discriminator', args=dict(ic_n=4, ch_base=32768, ch_max=512, resolution=res,
                                                                  use_fp16_before_res=None))).eval().requires_grad_(False).cuda()
    x, z = S.synthetic_batch(batch, res, seed=2)
    x, z = x.cuda(), z.cuda()

    def step():
        img = G(x, z, None, noise_mode='random')
        m = x[:, 0:1] + 0.5
        fake = x[:, 1:4] * m + img * (1 - m)                       # shgan_default.py:257-260 (float composite)
        return D(torch.cat([x[:, 0:1], fake], dim=1), None)
    ms = time_steps(step)
    ms_g = time_steps(lambda: G(x, z, None, noise_mode='random'))
    print(json.dumps(dict(metric='512x512 generator+discriminator forward steps, images/sec', value=batch / ms * 1e3, unit='images/s',
                          n_gpus=1, ms_per_step=ms, ms_generator_only=ms_g,
                          config=dict(workload='C4: Places2-512 shgan_places512_eval G forward -> composite -> D forward, batch 8 per GPU '
                                               '(forward only: the reference has no training step), synthetic inputs, random-init weights'),
                          whole_step_algorithmic_tflops=(GFLOP_G[res] + GFLOP_D512) * batch / ms)))


def c5():
    peak, src = hbm_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    rows = []
    ch = 32
    for r in (4, 8, 16, 32, 64, 128):
        n = max(16, min(4096, (256 << 20) // (ch * r * r * 4)))    # >= 256 MB of input where the batch limit allows
        lowest = 4
        masks = P.gaussian_band_masks(r, lowest, 3, False)
        reslist = sorted(masks)
        g = torch.Generator().manual_seed(r)
        c2_ = 2 * ch
        conv0_w = (torch.randn(c2_, c2_, generator=g) / 8).cuda()
        conv0_b = (torch.randn(c2_, generator=g) * 0.1).cuda()
        df1_w = (1 / 64 + 0.1 / 64 * torch.randn(c2_, c2_ * 6, generator=g)).cuda()
        cw = P.make_cweight((2, 3), (r, r // 2 + 1)).cuda().contiguous()
        gauss = torch.cat([masks[k].reshape(-1) for k in reslist]).cuda().contiguous()
        x = torch.randn(n, ch, r, r, device='cuda')
        outs = [torch.empty(n, ch, k, k, device='cuda') for k in reslist]
        ws = torch.empty(K.shu_workspace_bytes(n, ch, r), dtype=torch.uint8, device='cuda')

        def run():
            K.shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest, workspace=ws)
        for _ in range(3):
            run()
        ts = []
        for _ in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        byts = n * ch * 4 * (r * r + sum(k * k for k in reslist))
        bins = r * (r // 2 + 1)
        flops = n * bins * 2.0 * (c2_ * c2_ + c2_ * c2_ * 6)         # the two 1x1 channel mixes (complex = 2C real channels)
        rows.append(dict(input_res=r, batch=n, ms=ms, algorithmic_mb=byts / 1e6, gbs=byts / ms / 1e6, frac_of_hbm=byts / ms / 1e6 / peak,
                         channel_mix_gflops=flops / ms / 1e6))
    print(json.dumps(dict(metric='SHU rFFT2 + heterogeneous filter + Gaussian split + irFFT2, algorithmic GB/s', unit='GB/s',
                          config=dict(workload='C5: SHU-only sweep, 32 channels, lowest_res 4, input_res 4..128 (kernel range; the model uses 64)',
                                      l2_policy='256 MB buffer written between timed iterations'),
                          hbm_peak_gbs=peak, peak_source=src, sweep=rows)))


if __name__ == '__main__':
    which = [a.lower() for a in sys.argv[1:]] or ['c2', 'c4', 'c5']
    for w in which:
        {'c2': c2, 'c4': c4, 'c5': c5}[w]()
