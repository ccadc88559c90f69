"""Benchmark of the SH-GAN generator-forward hot path (BASELINE.json metric: 512x512 inpaint images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c5] [--impl shgan_b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Configurations (BASELINE.json `configs`):
  c3 (default)  FFHQ-512 `shgan_ffhq512_eval` generator forward, batch 16 per GPU, 1..8 GPUs batch-sharded  <- the headline
  c2            FFHQ-256 `shgan_ffhq256_eval` generator forward, batch 32, 1 GPU
  c4            Places2-512 generator + discriminator step (forward only: the reference ships no training step),
                batch 8 per GPU:  G(x, z) -> float composite + mask concat (one kernel) -> D
  c5            SHU-only sweep: rFFT2 + heterogeneous filter + Gaussian split + irFFT2 over input_res 4..512, GB/s vs HBM
One "step" = one pass of the configuration's hot path over one batch of synthetic free-form-masked images (masks =
the reference's RandomMask restated in shgan_b200/synthetic.py; random-init weights of the reference architecture).
Weak scaling: every rank runs its own batch, the forward needs no collective.  Rank 0 prints ONE JSON line:

  value          images/s with inputs resident in HBM, CUDA-event timed, max over ranks, CUDA-graph replay
  e2e            same metric through the public module API with pinned-host inputs: H2D copy of (x, z) and D2H read of the
                 uint8 composite inside the timed region; with N > 1 also a device-side feature reduction of every composite
                 and, at the end of the run, the eval loop's single NCCL all-gather of those features
  roofline       the dominant kernel family (tcgen05 implicit-GEMM convolutions): algorithmic FLOP/s from CUDA events around
                 every launch, against the measured bf16 tensor peak (MEASURED_PEAKS.json; the profiling guide's fallback
                 when the driver has not written the file -- `peak_source` says which)
  roofline_fir / roofline_fft
                 the HBM-bound kernels: the stand-alone blur passes of the step; the SHU at the model's size measured HBM-sized
                 (config C5's input_res-64 point, batch 512, L2 flushed; the L2-resident in-step call rides along as `in_step`):
                 algorithmic GB/s against the measured HBM copy peak
  roofline_upfirdn2d
                 the op-level `shgan_upfirdn2d_fwd` (1:1 with the reference plugin's entry point) on the model's blur / up / down
                 variants: algorithmic GB/s against the HBM peak, the reference's CUDA plugin timed beside it (N = 1 only)
  collective     (N > 1) the all-gather of [items, 2*2048+1] float64 detector features: bytes, ms, GB/s, order checked
  reference_gpu  the UNMODIFIED reference generator (baseline/_ref: cuDNN fp32, TF32 off, its own upfirdn2d CUDA plugin)
                 timed with the same CUDA events on the same GPU, same batch -- the same-box GPU speed-up column
  other_configs  c2 / c4 / c5 summaries measured in the same run (N = 1 only; `--no-other-configs` skips them)
  cpu_baseline   the reference's own CPU path on the host cores, bounded sample
`--impl reference` is the reference arm: the unmodified reference generator (baseline/_ref) on the host cores
(`kind: "reference"`; the oracle port only if the copy is absent, `kind: "port"`), K steps after W warm-ups, each step a
bounded sample of the configuration's batch.  `--impl reference --ref-device cuda` times it on the GPU instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

GFLOP_G = {512: 238.785, 256: 180.635}      # SURVEY.md section 8d (algorithmic, 2*MAC)
GFLOP_D512 = 123.204
SHU_BYTES_PER_SAMPLE = 1222656              # 524 288 in + 698 368 out at input_res 64, 32 channels
FEATURE_DIM = 2048

CONFIGS = {
    'c2': dict(res=256, batch=32, kind='gen', metric='256x256 inpaint images/sec',
               workload='FFHQ-256 shgan_ffhq256_eval generator forward, batch 32/GPU, synthetic free-form masks, random-init weights'),
    'c3': dict(res=512, batch=16, kind='gen', metric='512x512 inpaint images/sec',
               workload='FFHQ-512 shgan_ffhq512_eval generator forward, batch 16/GPU, synthetic free-form masks, random-init weights'),
    'c4': dict(res=512, batch=8, kind='gen+disc', metric='512x512 generator+discriminator forward images/sec',
               workload='Places2-512 shgan_places512_eval G forward -> composite+concat -> D forward, batch 8/GPU (forward only: the '
                        'reference has no training step), synthetic free-form masks, random-init weights'),
    'c5': dict(res=64, batch=0, kind='shu', metric='SHU rFFT2 + heterogeneous filter + Gaussian split + irFFT2, algorithmic GB/s',
               workload='SHU-only sweep, 32 channels, lowest_res 4, input_res 4..512'),
}


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tensor=d['bf16_tflops_sustained'], tensor_burst=d['bf16_tflops'], src='measured')
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '200'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out['sm_max_mhz'] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[5:9]):
                if v.strip().lower().startswith('active'):
                    reasons.add(name)
        if sm:
            busy = [v for v in sm if v >= 0.6 * max(sm)] or sm           # samples taken under load
            out['sm_mhz'] = statistics.median(busy)
        out['reasons'] = sorted(reasons)
        out['samples'] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def _reference_generator(res, device):
    """The unmodified reference generator with the benchmark's random-init weights (same state_dict as the repo arm), or
    None when no reference tree is available (then the oracle port is the stand-in)."""
    import torch
    from golden import ref_import
    if not ref_import.reference_available():
        return None, None
    from shgan_b200 import synthetic as S
    torch.backends.cuda.matmul.allow_tf32 = False                  # configs/experiment/shgan_ffhq512_eval.yaml:10-11
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    R = ref_import.import_reference()
    Gr = ref_import.build_reference_generator(R, res)
    Gr.load_state_dict(S.random_generator(res, seed=0, device='cpu').state_dict(), strict=True)
    return Gr.to(device), R


def reference_cpu_run(cfg, steps, warmup, budget_s=150.0):
    """The reference's own CPU path (PyTorch CPU convs through conv2d_gradfix.py:38,43 + the pure-torch upfirdn2d,
    upfirdn2d.py:98-138) on all host threads.  Every step processes the same bounded sample of the configuration's batch;
    the sample size is chosen from a first timed image so that warm-up + steps fit in `budget_s`."""
    import numpy as np
    import torch
    from shgan_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count() or 1)
    res = cfg['res']
    x, z = S.synthetic_batch(cfg['batch'], res, seed=1000)
    Gr, _ = _reference_generator(res, 'cpu')
    if Gr is not None:
        kind, what = 'reference', 'unmodified reference comodgan.Generator (baseline/_ref), PyTorch CPU backend'

        def run(n):
            with torch.no_grad():
                return Gr(x[:n], z[:n], torch.zeros(n, 0), noise_mode='random')
    else:
        from oracle import shgan_oracle as O
        O.set_conv_backend('torch')
        sd = {k: v.numpy() for k, v in S.random_generator(res, seed=0, device='cpu').state_dict().items()}
        kind, what = 'port', 'oracle/shgan_oracle.py port of the reference path (torch CPU conv backend); baseline/_ref absent'

        def run(n):
            return O.generator(sd, x[:n].numpy(), z[:n].numpy(), res)
    t0 = time.perf_counter()
    run(1)                                                         # also the first warm-up (allocations, thread pools)
    t1 = time.perf_counter() - t0
    n = int(max(1, min(cfg['batch'], budget_s / max(t1, 1e-3) / max(steps + warmup, 1))))
    for _ in range(warmup):
        run(n)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(n)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    cb = dict(value=n / sec, unit='images/s', cores=torch.get_num_threads(), kind=kind,
              sample=f'{steps} steps x {n} of the {cfg["batch"]} images of a batch ({res}x{res} generator forward, noise_mode=random) after '
                     f'{warmup} warm-up steps; {what}')
    return cb, sec, n


def reference_gpu_run(cfg, dev, x_d, z_d, steps=5, warmup=3):
    """The unmodified reference on the same GPU: cuDNN fp32 (TF32 off) + the reference's upfirdn2d CUDA plugin."""
    import torch
    Gr, R = _reference_generator(cfg['res'], dev)
    if Gr is None:
        return dict(unavailable='baseline/_ref absent (run baseline/install_reference.py in the build container)')
    c = torch.zeros(x_d.shape[0], 0, device=dev)
    with torch.no_grad():
        for _ in range(warmup):
            Gr(x_d, z_d, c, noise_mode='random')
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            Gr(x_d, z_d, c, noise_mode='random')
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    plugin = 'cuda plugin (upfirdn2d.cu, JIT-built)' if R.upfirdn2d._init() else 'pure-torch fallback (plugin build failed)'
    del Gr
    torch.cuda.empty_cache()
    return dict(value=x_d.shape[0] / ms * 1e3, unit='images/s', ms_per_step=ms, steps=steps, warmup=warmup, batch=int(x_d.shape[0]),
                what=f'unmodified reference comodgan.Generator.forward(noise_mode=random) on this GPU: cuDNN fp32, allow_tf32=False, '
                     f'cudnn.benchmark=False (configs/experiment/shgan_ffhq512_eval.yaml:10-11), upfirdn2d = {plugin}')


# ------------------------------------------------------------------------------------------------ SHU sweep (c5)
def shu_sweep(dev, peaks, resolutions=(4, 8, 16, 32, 64, 128, 256, 512), iters=11):
    import numpy as np
    import torch
    from shgan_b200 import kernels as K, packing as P
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    ch = 32
    for r in resolutions:
        n = max(4, min(4096, (256 << 20) // (ch * r * r * 4)))    # >= 256 MB of input where the batch limit allows
        lowest = 4
        masks = P.gaussian_band_masks(r, lowest, 3, False)
        reslist = sorted(masks)
        g = torch.Generator().manual_seed(r)
        c2_ = 2 * ch
        conv0_w = (torch.randn(c2_, c2_, generator=g) / 8).to(dev)
        conv0_b = (torch.randn(c2_, generator=g) * 0.1).to(dev)
        df1_w = (1 / 64 + 0.1 / 64 * torch.randn(c2_, c2_ * 6, generator=g)).to(dev)
        cw = P.make_cweight((2, 3), (r, r // 2 + 1)).to(dev).contiguous()
        gauss = torch.cat([masks[k].reshape(-1) for k in reslist]).to(dev).contiguous()
        x = torch.randn(n, ch, r, r, device=dev)
        outs = [torch.empty(n, ch, k, k, device=dev) for k in reslist]
        try:
            ws = torch.empty(K.shu_workspace_bytes(n, ch, r), dtype=torch.uint8, device=dev)
            packed = K.shu_pack(conv0_w, df1_w)              # once per parameter set, as engine.refresh does

            def run():
                K.shu_fwd(x, conv0_w, conv0_b, df1_w, cw, gauss, outs, lowest, workspace=ws, packed=packed)
            for _ in range(3):
                run()
            for _ in range(3):                               # and through the flush pattern of the timed loop (fresh pages, TLB, clocks)
                flush.zero_()
                run()
        except RuntimeError as ex:
            rows.append(dict(input_res=r, batch=n, unsupported=str(ex)[:120]))
            continue
        ts = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        byts = n * ch * 4 * (r * r + sum(k * k for k in reslist))
        bins = r * (r // 2 + 1)
        flops = n * bins * 2.0 * (c2_ * c2_ + c2_ * c2_ * 6)         # the two 1x1 channel mixes (complex = 2C real channels)
        rows.append(dict(input_res=r, batch=n, ms=ms, algorithmic_mb=byts / 1e6, gbs=byts / ms / 1e6,
                         frac_of_hbm=byts / ms / 1e6 / peaks['hbm'], channel_mix_tflops=flops / ms / 1e9))
        del x, outs, ws
    del flush
    torch.cuda.empty_cache()
    return rows


# ------------------------------------------------------------------------------------------------ upfirdn2d (the plugin's op)
def upfirdn2d_roofline(dev, peaks, iters=7):
    """`shgan_upfirdn2d_fwd` (the C-ABI entry point that replaces the reference's only native code, upfirdn2d.cpp:16 /
    upfirdn2d.cu:29-200) through `ops.upfirdn2d`, on the variants the model runs at its largest resolution, batch 16:
    algorithmic in + out bytes / CUDA-event time against the HBM peak, L2 flushed between iterations; the reference's own CUDA
    plugin is timed on the same tensors when baseline/_ref is present."""
    import numpy as np
    import torch
    from shgan_b200 import ops
    from golden import ref_import
    R = None
    if ref_import.reference_available():
        try:
            R = ref_import.import_reference()
            if not R.upfirdn2d._init():
                R = None
        except Exception:
            R = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fo = ops.setup_filter([1, 3, 3, 1], device=torch.device(dev))
    fr = R.upfirdn2d.setup_filter([1, 3, 3, 1], device=torch.device(dev)) if R is not None else None

    def med(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts))
    cases = [('blur before the stride-2 conv (conv2d_resample.py:117-120): pad 2', (16, 64, 512, 512), dict(padding=[2, 2, 2, 2])),
             ('blur after the transposed conv (conv2d_resample.py:139): pad 1, gain 4', (16, 64, 513, 513), dict(padding=[1, 1, 1, 1], gain=4)),
             ('upsample2d of the image (comodgan.py:331-338): up 2', (16, 3, 256, 256), dict(up=2, padding=[2, 1, 2, 1], gain=4)),
             ('downsample2d (stylegan.py:658-684 skip path): down 2', (16, 64, 512, 512), dict(down=2, padding=[1, 1, 1, 1]))]
    rows = []
    for what, shape, kw in cases:
        x = torch.randn(shape, device=dev)
        y = ops.upfirdn2d(x, fo, **kw)
        byts = 4.0 * (x.numel() + y.numel())
        ms = med(lambda: ops.upfirdn2d(x, fo, **kw))
        row = dict(case=what, shape=list(shape), ms=ms, algorithmic_mb=byts / 1e6, gbs=byts / ms / 1e6, frac_of_hbm=byts / ms / 1e6 / peaks['hbm'])
        if R is not None:
            with torch.no_grad():
                ms_r = med(lambda: R.upfirdn2d.upfirdn2d(x, fr, impl='cuda', **kw))
            row['reference_plugin_ms'] = ms_r
            row['reference_plugin_gbs'] = byts / ms_r / 1e6
        rows.append(row)
        del x, y
    del flush
    torch.cuda.empty_cache()
    best = rows[0]
    return dict(bound='hbm', kernel='shgan_upfirdn2d_fwd: upfirdn2d_tile_kernel (stride-1 blurs) / upfirdn2d_gather_kernel (up / down variants), NCHW fp32',
                achieved=best['gbs'], peak=peaks['hbm'], unit='GB/s', frac=best['frac_of_hbm'], traffic=None,
                peak_source=f'{peaks["src"]} HBM copy', l2_policy='256 MB buffer written between timed iterations',
                note='headline = the pad-2 blur of a [16,64,512,512] tensor; every case of the model in `cases`', cases=rows)


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='shgan_b200', choices=['shgan_b200', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--res', type=int, default=None, help='override the configuration resolution')
    ap.add_argument('--batch', type=int, default=None, help='override images per GPU per step')
    ap.add_argument('--passes', type=int, default=3, help='3 = fp32-class split-precision convs (parity mode), 1 = single fp16 pass')
    ap.add_argument('--ref-device', default='cpu', choices=['cpu', 'cuda'], help='device of the --impl reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-gpu', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = dict(CONFIGS[args.config])
    if args.res:
        cfg['res'] = args.res
    if args.batch:
        cfg['batch'] = args.batch
    res, batch = cfg['res'], cfg['batch']
    warmup = max(args.warmup, 3)

    def config_block(extra=None):
        c = dict(workload=cfg['workload'], config=args.config, resolution=res, batch_per_gpu=batch, global_batch=batch * world,
                 parallelism=f'dp{world} (batch-sharded, no collective in the forward)', noise_mode='random',
                 masks='RandomMask(R, hole_range=[0,1]) restated from lib/data_factory/ds_ffhq.py:145-217',
                 precision='fp16 hi+lo split operands x3 tensor-core passes, fp32 accumulation (fp32-class, 1e-3 max-abs parity)'
                 if args.passes == 3 else 'single fp16 pass (NOT parity mode)',
                 l2_policy='working set per step (>1 GB of activations) exceeds the 126 MB L2; no explicit flush',
                 gflop_per_image=GFLOP_G.get(res))
        c.update(extra or {})
        return c

    # ---------------------------------------------------------------- the reference arm
    if args.impl == 'reference':
        if rank != 0:
            return
        if cfg['kind'] == 'shu':
            print(json.dumps(dict(impl='reference', unavailable='the SHU sweep has no reference arm; use --config c3')))
            return
        if args.ref_device == 'cuda':
            import torch
            from shgan_b200 import synthetic as S
            dev = torch.device('cuda', local_rank)
            x, z = S.synthetic_batch(batch, res, seed=1000)
            r = reference_gpu_run(cfg, dev, x.to(dev), z.to(dev), steps=args.steps, warmup=warmup)
            line = dict(metric=cfg['metric'], value=r.get('value'), unit='images/s', n_gpus=1, steps=args.steps, warmup=warmup,
                        ms_per_step=r.get('ms_per_step'), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                        data='synthetic', impl='reference', config=config_block(dict(cuda_graph=not args.no_graphs)),
                        reference_note=r.get('what', r.get('unavailable')),
                        e2e=dict(value=r.get('value'), unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
            print(json.dumps(line))
            return
        cb, sec, n = reference_cpu_run(cfg, args.steps, args.warmup)
        line = dict(metric=cfg['metric'], value=cb['value'], unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                    impl='reference', config=config_block(dict(cuda_graph=not args.no_graphs)),
                    reference_note='the config block is the repo arm\'s (same workload); this arm ran it on the host cores only, see cpu_baseline.sample',
                    cpu_baseline=cb, e2e=dict(value=cb['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- the repo arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl shgan_b200 needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from shgan_b200 import _lib, kernels as K, synthetic as S, parallel as PL
    from shgan_b200.model_zoo import get_model
    dev = torch.device('cuda', local_rank)
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if cfg['kind'] == 'shu':
        if rank == 0:
            rows = shu_sweep(dev, peaks)
            best = max((r for r in rows if 'gbs' in r), key=lambda r: r['gbs'])
            at64 = [r for r in rows if r.get('input_res') == 64 and 'gbs' in r]
            print(json.dumps(dict(metric=cfg['metric'], value=(at64[0] if at64 else best)['gbs'], unit='GB/s', n_gpus=1, steps=7, warmup=3,
                                  higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                                  config=dict(workload=cfg['workload'], config='c5', l2_policy='256 MB buffer written between timed iterations',
                                              value_is='GB/s at input_res 64 (the released model\'s SHU size)'),
                                  roofline=dict(bound='hbm', achieved=(at64[0] if at64 else best)['gbs'], peak=peaks['hbm'], unit='GB/s',
                                                frac=(at64[0] if at64 else best)['gbs'] / peaks['hbm'], traffic=None,
                                                peak_source=f'{peaks["src"]} HBM copy'),
                                  sweep=rows)))
        if world > 1:
            dist.destroy_process_group()
        return

    G = S.random_generator(res, seed=0, device=dev)
    eng = G.engine(passes=args.passes, impl=0)
    D = None
    if cfg['kind'] == 'gen+disc':
        torch.manual_seed(3)
        D = get_model()(dict(type='comodgan_discriminator', args=dict(ic_n=4, ch_base=32768, ch_max=512, resolution=res,
                                                                      use_fp16_before_res=None))).eval().requires_grad_(False).to(dev)
    x_h, z_h = S.synthetic_batch(batch, res, seed=1000 + rank)
    x_pin, z_pin = x_h.pin_memory(), z_h.pin_memory()
    x_d, z_d = x_pin.to(dev), z_pin.to(dev)
    comp_pin = torch.empty((batch, 3, res, res), dtype=torch.uint8).pin_memory()

    # ---- instrumentation: CUDA events around every launch of the three kernel families (current torch stream) ----
    fam = {k: dict(events=[], work=0.0) for k in ('conv', 'up2', 'fir', 'shu')}
    record = [False]
    orig = dict(conv=K.conv_igemm, up2=K.conv_up2, fir=K.fir_nhwc, shu=K.shu_fwd)

    def timed(name, work_fn):
        fn = orig[name]

        def wrapped(*a, **kw):
            if not record[0]:
                return fn(*a, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **kw)
            e.record()
            fam[name]['events'].append((s, e))
            fam[name]['work'] += work_fn(*a, **kw)
            return out
        return wrapped

    def conv_work(srcs, w_hi, w_lo, taps, oh, ow, **kw):
        n, _, _, c = srcs[0].shape
        return 2.0 * n * oh * ow * w_hi.shape[1] * c * len(taps)

    def up2_work(src, w_hi, *a, **kw):
        n, h, w, c = src.shape
        return 2.0 * n * h * w * 9 * c * w_hi.shape[0] * 64            # transposed conv counted on its input grid

    def fir_work(src, f, gain, pads, epi, parity_split=False, rank1=False):
        n, ih, iw, c = src.shape
        oh, ow = ih + pads[2] + pads[3] - 3, iw + pads[0] + pads[1] - 3
        if parity_split == 2:
            oh, ow = (oh + 1) // 2, (ow + 1) // 2
        return 4.0 * n * c * (ih * iw + oh * ow)                    # fp32-equivalent bytes in + out

    def shu_work(x, *a, **kw):
        outs = a[5] if len(a) > 5 else kw['outs']
        return 4.0 * (x.numel() + sum(o.numel() for o in outs))
    K.conv_igemm, K.fir_nhwc, K.shu_fwd = timed('conv', conv_work), timed('fir', fir_work), timed('shu', shu_work)
    K.conv_up2 = timed('up2', up2_work)

    if D is None:
        def step_device():
            return G.forward_composite(x_d, z_d, noise_mode='random')
    else:
        def step_device():
            img = G.engine().forward(x_d, z_d, noise_mode='random')
            return D(K.composite_cat(x_d, img), None)

    # end-to-end: what an eval loop does with this API -- pinned host batches in, uint8 composites out -- as a 2-deep
    # pipeline: the H2D copy of batch k+1 and the D2H copy of result k-1 run on a copy stream while batch k computes.
    copy_s = torch.cuda.Stream()
    dbuf = [(torch.empty_like(x_d), torch.empty_like(z_d)) for _ in range(2)]
    comp_dev = [torch.empty((batch, 3, res, res), dtype=torch.uint8, device=dev) for _ in range(2)]
    comp_pins = [comp_pin, torch.empty_like(comp_pin).pin_memory()]
    gather_info = {}

    def detector(u8):
        """Synthetic stand-in for the Inception detector (its TorchScript is fetched from a CDN, unreachable offline,
        lib/evaluator/eva_fid.py:21): [b, 2048] float64 block means of the uint8 composite, computed on the device."""
        b = u8.shape[0]
        return u8.reshape(b, FEATURE_DIM, -1).float().mean(-1).to(torch.float64)

    def run_e2e(k_steps):
        main_s = torch.cuda.current_stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        d2h = [torch.cuda.Event() for _ in range(2)]
        feats = []
        with torch.cuda.stream(copy_s):
            dbuf[0][0].copy_(x_pin, non_blocking=True)
            dbuf[0][1].copy_(z_pin, non_blocking=True)
            ready[0].record(copy_s)
        for k in range(k_steps):
            cur, nxt = k & 1, (k + 1) & 1
            if k + 1 < k_steps:
                with torch.cuda.stream(copy_s):
                    if k >= 1:
                        copy_s.wait_event(freed[nxt])
                    dbuf[nxt][0].copy_(x_pin, non_blocking=True)
                    dbuf[nxt][1].copy_(z_pin, non_blocking=True)
                    ready[nxt].record(copy_s)
            main_s.wait_event(ready[cur])
            if k >= 2:
                main_s.wait_event(d2h[cur])                      # comp_dev[cur] has been read back
            _, comp = G.forward_composite(dbuf[cur][0], dbuf[cur][1], noise_mode='random')
            comp_dev[cur].copy_(comp, non_blocking=True)       # the engine's output buffer is reused by the next replay
            if world > 1:
                feats.append(detector(comp_dev[cur]))
            freed[cur].record(main_s)
            done[cur].record(main_s)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done[cur])
                comp_pins[cur].copy_(comp_dev[cur], non_blocking=True)
                d2h[cur].record(copy_s)
        if world > 1:
            # the eval loop's ONLY exchange: one all-gather of every rank's features at the end of the run
            local = torch.cat(feats)
            n_items = local.shape[0] * world
            full = PL.gather_features(local, n_items)
            gather_info['e2e_bytes'] = int(full.numel() * 8)
        copy_s.synchronize()
        main_s.synchronize()

    # one eager (un-graphed) step counts the kernels of a step; the product path then replays them as a CUDA graph
    eng.graphs = False
    step_device()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    step_device()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0
    eng.graphs = not args.no_graphs
    for _ in range(warmup):
        step_device()
    barrier()

    # ---- device-resident timing (product path) ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = launches_per_step * args.steps

    # ---- same K steps again, launched eagerly with CUDA events around every launch of the three kernel families (events
    #      cannot be read back from inside a replayed graph) ---------------------------------------------------------
    eng.graphs = False
    eng.overlap = False           # single stream: the events then bracket exactly one kernel family's launches
    record[0] = True
    barrier()
    evi0, evi1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evi0.record()
    for _ in range(args.steps):
        step_device()
    evi1.record()
    barrier()
    record[0] = False
    ms_instr = evi0.elapsed_time(evi1)
    fam_ms = {k: sum(s.elapsed_time(e) for s, e in v['events']) for k, v in fam.items()}
    eng.graphs = not args.no_graphs
    eng.overlap = True

    # ---- end-to-end timing (pinned host inputs, uint8 composite read back) ------------------------------------------
    e2e_ms = None
    if D is None:
        run_e2e(3)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)                                   # ends with both streams synchronised: wall clock == device time
        e2e_ms = (time.perf_counter() - t0) * 1e3
        barrier()
    clocks = sampler.stop() if sampler else None

    # ---- the eval loop's collective on its own (N > 1): EvalLoop over a synthetic dataset, ONE all-gather, order checked ----
    collective = None
    if world > 1 and D is None:
        n_items = world * batch * 2 + 3                         # not a multiple of the world size: exercises the wrap-around
        rs_cache = {}

        def dataset(i):
            g = torch.Generator().manual_seed(10_000 + i)
            real = torch.randn(3, res, res, generator=g).clamp_(-1, 1)
            if i not in rs_cache:
                import numpy as np
                rs_cache[i] = torch.from_numpy(S.freeform_mask(res, np.random.RandomState(i)))
            return real, rs_cache[i]
        loop = PL.EvalLoop(G, detector, batch, dev, z_seed=1)
        gather_ms = []
        orig_gather = PL.gather_features

        def gather_timed(local, n, group=None):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            s.record()
            out = orig_gather(local, n, group)
            e.record()
            torch.cuda.synchronize()
            gather_ms.append(s.elapsed_time(e))
            return out
        PL.gather_features = gather_timed
        fake, real = loop.run(dataset, n_items)                 # raises if the gathered order is not the dataset order
        fake, real = loop.run(dataset, n_items)                 # second run: NCCL communicator and buffers are warm
        PL.gather_features = orig_gather
        tg = torch.tensor([gather_ms[-1]], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gms = float(tg[0])
        collective = dict(op='all_gather (NCCL), once per eval run', items=n_items, feature_dim=2 * FEATURE_DIM + 1, dtype='f64',
                          bytes=loop.last_gather_bytes, ms=gms, gbs=loop.last_gather_bytes / gms / 1e6,
                          order_checked='gathered index column == arange(items) on the device; wrap-around duplicates dropped',
                          first_call_ms=gather_ms[0], fid_of_synthetic_features=PL.fid_from_features(fake.cpu().numpy(), real.cpu().numpy()))

    vals = [ms, e2e_ms if e2e_ms is not None else 0.0]
    tms = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), (float(tms[1]) if e2e_ms is not None else None)

    if rank == 0:
        imgs = batch * world * args.steps
        value = imgs / (ms / 1e3)
        steps = max(args.steps, 1)
        # the convolution family = shgan_conv_igemm + the fused up-sampling convolution shgan_conv_up2
        fam['conv']['work'] += fam['up2']['work']
        fam['conv']['events'] += fam['up2']['events']
        fam_ms['conv'] += fam_ms['up2']
        conv_tflops = fam['conv']['work'] / (fam_ms['conv'] / 1e3) / 1e12 if fam_ms['conv'] > 0 else 0.0
        tp = os.path.join(ROOT, 'profiles', 'conv_tc_traffic.json')
        traffic, traffic_src = None, None
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get('dram_bytes_per_launch_avg'), tj.get('source')
        gflop_step = (GFLOP_G.get(res, 0) + (GFLOP_D512 if D is not None and res == 512 else 0)) * batch * world
        line = dict(
            metric=cfg['metric'], value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=warmup,
            ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f32' if args.passes == 3 else 'f16', data='synthetic',
            config=config_block(dict(cuda_graph=not args.no_graphs)),
            clocks=clocks,
            gpu_launches=int(launches),
            roofline=dict(bound='tensor', kernel='shgan_conv_igemm: conv_pair_kernel (tcgen05 cta_group::2) + conv_up2_kernel (fused transposed conv + blur) '
                                                 '+ conv_tc_kernel / conv_halo_kernel (single-CTA tcgen05), every conv launch of a step',
                          achieved=conv_tflops, peak=peaks['tensor'], unit='TFLOP/s', frac=conv_tflops / peaks['tensor'],
                          traffic=traffic, traffic_source=traffic_src or 'committed ncu capture (profiles/), not re-measured in this run',
                          peak_source=f'{peaks["src"]} bf16 sustained',
                          launches_per_step=len(fam['conv']['events']) // steps,
                          algorithmic_gflop_per_step=fam['conv']['work'] / steps / 1e9,
                          kernel_ms_per_step=fam_ms['conv'] / steps, share_of_step=fam_ms['conv'] / ms_instr,
                          up2_ms_per_step=fam_ms['up2'] / steps, up2_launches_per_step=len(fam['up2']['events']) // steps,
                          instrumented_ms_per_step=ms_instr / steps,
                          executed_tensor_tflops=conv_tflops * (3 if args.passes == 3 else 1),
                          note='achieved = algorithmic conv FLOPs / summed CUDA-event durations of the conv launches over the same K steps launched eagerly on one stream '
                               '(value/ms_per_step come from the CUDA-graph replay of the identical launch sequence); '
                               'parity mode issues 3 fp16 MMA passes per algorithmic FLOP, so frac <= 1/3 by construction'),
            whole_step_algorithmic_tflops=gflop_step / (ms / args.steps) if gflop_step else None,
        )
        if fam_ms['fir'] > 0:
            gbs = fam['fir']['work'] / fam_ms['fir'] / 1e6
            line['roofline_fir'] = dict(bound='hbm', kernel='shgan_fir_nhwc (blur in front of the stride-2 convs: fir4x4_walk_kernel)',
                                        achieved=gbs, peak=peaks['hbm'], unit='GB/s', frac=gbs / peaks['hbm'], traffic=None,
                                        peak_source=f'{peaks["src"]} HBM copy', launches_per_step=len(fam['fir']['events']) // steps,
                                        kernel_ms_per_step=fam_ms['fir'] / steps, algorithmic_mb_per_step=fam['fir']['work'] / steps / 1e6)
        if fam_ms['shu'] > 0:
            gbs = fam['shu']['work'] / fam_ms['shu'] / 1e6
            in_step = dict(achieved=gbs, frac=gbs / peaks['hbm'], kernel_ms_per_step=fam_ms['shu'] / steps,
                           algorithmic_mb_per_step=fam['shu']['work'] / steps / 1e6,
                           note=f'batch {batch}: 20 MB working set, L2-resident and latency-bound, on the side stream of the step')
            line['roofline_fft'] = dict(bound='hbm', kernel='shgan_shu_fwd (rFFT2 + channel mix + heterogeneous filter + Gaussian split + 5 irFFT2), inside the step',
                                        unit='GB/s', peak=peaks['hbm'], peak_source=f'{peaks["src"]} HBM copy', traffic=None, **in_step)
            if world == 1:
                # the HBM-sized measurement of the same entry point at the model's size (config C5's input_res-64 point: batch 512,
                # 268 MB in, 358 MB out, L2 flushed between iterations) is the roofline figure; the in-step number rides along
                try:
                    row = shu_sweep(dev, peaks, resolutions=(64,))[0]
                    line['roofline_fft'] = dict(
                        bound='hbm', kernel='shgan_shu_fwd at the released model\'s size (input_res 64, 32 channels), batch 512: shu_rfft2_r64_kernel + '
                                            'shu_mix_tc_kernel (tcgen05) + shu_irfft2_r64_kernel',
                        achieved=row['gbs'], peak=peaks['hbm'], unit='GB/s', frac=row['frac_of_hbm'], peak_source=f'{peaks["src"]} HBM copy',
                        traffic=1593.6e6, traffic_source='profiles/r2_shu_ncu_summary.md (ncu --set full, dram bytes read + written by the three '
                                                         'launches of one call; a cited ncu figure: the spectrum crosses HBM twice)',
                        ms_per_call=row['ms'], algorithmic_mb_per_call=row['algorithmic_mb'], l2_policy='256 MB buffer written between timed iterations',
                        in_step=in_step)
                except Exception as ex:
                    line['roofline_fft']['c5_point_error'] = f'{type(ex).__name__}: {str(ex)[:200]}'

        if e2e_ms is not None:
            line['e2e'] = dict(value=imgs / (e2e_ms / 1e3), unit='images/s', ms_per_step=e2e_ms / args.steps,
                               h2d_bytes_per_step=int(x_pin.numel() * 4 + z_pin.numel() * 4), d2h_bytes_per_step=int(comp_pin.numel()),
                               api='model_zoo.comodgan.Generator.forward_composite(x, z); pinned host batches in, uint8 composites read back to pinned host; '
                                   'H2D/D2H on a copy stream, 2-deep pipeline, all copies inside the timed region'
                                   + ('; plus per-step device-side features and ONE NCCL all-gather of them at the end' if world > 1 else ''),
                               gather_bytes=gather_info.get('e2e_bytes'))
        if collective is not None:
            line['collective'] = collective
        if world == 1:
            if not args.no_reference_gpu and D is None:
                try:
                    rg = reference_gpu_run(cfg, dev, x_d, z_d)
                except Exception as ex:                          # plugin JIT / import problems must not lose the bench line
                    rg = dict(unavailable=f'{type(ex).__name__}: {str(ex)[:200]}')
                if 'value' in rg:
                    rg['speedup_device_resident'] = value / rg['value']
                line['reference_gpu'] = rg
            if not args.no_other_configs and args.config == 'c3':
                try:
                    line['roofline_upfirdn2d'] = upfirdn2d_roofline(dev, peaks)
                except Exception as ex:
                    line['roofline_upfirdn2d'] = dict(unavailable=f'{type(ex).__name__}: {str(ex)[:200]}')
                line['other_configs'] = other_configs(dev, peaks)
            if not args.no_cpu_baseline and D is None:
                cb, _, _ = reference_cpu_run(cfg, 2, 1, budget_s=25.0)
                line['cpu_baseline'] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(dev, peaks):
    """c2 / c4 / c5 in the same process (device-resident timing only), so that the driver's default run sees every
    BASELINE.json configuration with the measured peaks."""
    import torch
    from shgan_b200 import kernels as K, synthetic as S
    from shgan_b200.model_zoo import get_model

    def time_steps(fn, steps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / steps
    out = {}
    try:
        c = CONFIGS['c2']
        G = S.random_generator(c['res'], seed=0, device=dev)
        x, z = S.synthetic_batch(c['batch'], c['res'], seed=1)
        x, z = x.to(dev), z.to(dev)
        ms = time_steps(lambda: G.forward_composite(x, z, noise_mode='random'))
        out['c2'] = dict(metric=c['metric'], value=c['batch'] / ms * 1e3, unit='images/s', ms_per_step=ms, workload=c['workload'],
                         whole_step_algorithmic_tflops=GFLOP_G[c['res']] * c['batch'] / ms)
        del G, x, z
        torch.cuda.empty_cache()
        c = CONFIGS['c4']
        G = S.random_generator(c['res'], seed=0, device=dev)
        torch.manual_seed(3)
        D = get_model()(dict(type='comodgan_discriminator', args=dict(ic_n=4, ch_base=32768, ch_max=512, resolution=c['res'],
                                                                      use_fp16_before_res=None))).eval().requires_grad_(False).to(dev)
        x, z = S.synthetic_batch(c['batch'], c['res'], seed=2)
        x, z = x.to(dev), z.to(dev)

        def step():
            img = G.engine().forward(x, z, noise_mode='random')
            return D(K.composite_cat(x, img), None)
        ms = time_steps(step)
        ms_g = time_steps(lambda: G.engine().forward(x, z, noise_mode='random'))
        out['c4'] = dict(metric=c['metric'], value=c['batch'] / ms * 1e3, unit='images/s', ms_per_step=ms, ms_generator_only=ms_g,
                         workload=c['workload'], whole_step_algorithmic_tflops=(GFLOP_G[c['res']] + GFLOP_D512) * c['batch'] / ms,
                         note='1 GPU, batch 8; the 8-GPU point of this config is `bench.py --config c4 --gpus 8`')
        del G, D, x, z
        torch.cuda.empty_cache()
        rows = shu_sweep(dev, peaks)
        out['c5'] = dict(metric=CONFIGS['c5']['metric'], unit='GB/s', hbm_peak_gbs=peaks['hbm'], peak_source=peaks['src'],
                         l2_policy='256 MB buffer written between timed iterations', sweep=rows)
    except Exception as ex:
        out['error'] = f'{type(ex).__name__}: {str(ex)[:300]}'
    return out


if __name__ == '__main__':
    main()
