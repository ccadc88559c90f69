"""Benchmark of the SH-GAN generator-forward hot path (BASELINE.json metric: 512x512 inpaint images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl shgan_b200|reference] [--res 512] [--batch 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one generator forward (mapping + encoder + SHU + synthesis) over one batch of synthetic free-form-masked
images of the FFHQ-512 config (`shgan_ffhq512_eval`, batch 16 per GPU; weak scaling: every rank runs its own batch,
the forward needs no collective).  Rank 0 prints ONE JSON line:
  value        images/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public module API with pinned-host inputs: H2D copy of (x, z) and D2H read of the
               uint8 composite inside the timed region
  roofline     the dominant kernel (the tcgen05 implicit-GEMM convolution behind shgan_conv_igemm): algorithmic FLOP/s measured
               with CUDA events around every launch, against the measured bf16 tensor peak of MEASURED_PEAKS.json (the
               profiling guide's fallback when the driver has not written that file; `peak_source` says which)
  cpu_baseline the CPU oracle (a port of the reference's PyTorch CPU path; the reference is Python and cannot be compiled
               into oracle/_ref) timed on the host cores on a bounded sample
`--impl reference` times that CPU implementation alone, as the reference arm.
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

GFLOP_PER_IMAGE = {512: 238.785, 256: 180.635}      # SURVEY.md section 8d (algorithmic, 2*MAC)
METRIC = '512x512 inpaint images/sec'


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
            out['sm_mhz'] = statistics.median(sm)
        out['reasons'] = sorted(reasons)
        out['samples'] = len(sm)
        return out


def cpu_reference_run(res, steps, warmup):
    """The reference's CPU path restated by the oracle (conv backend = torch.nn.functional.conv2d on all host threads,
    the call the reference itself makes on CPU through conv2d_gradfix.py:38,43).  One step = one image."""
    import numpy as np
    import torch
    from oracle import shgan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    O.set_conv_backend('torch')
    sd = O.synthetic_state_dict(res, seed=0)
    x, z = O.synthetic_inputs(1, res, seed=0)
    for _ in range(warmup):
        O.generator(sd, x, z, res)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.generator(sd, x, z, res)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return dict(value=1.0 / sec, unit='images/s', cores=torch.get_num_threads(), kind='port',
                sample=f'{steps} x 1 image {res}x{res} generator forward after {warmup} warm-up (oracle/shgan_oracle.py, torch CPU conv backend)'), sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='shgan_b200', choices=['shgan_b200', 'reference'])
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--batch', type=int, default=16, help='images per GPU per step')
    ap.add_argument('--passes', type=int, default=3, help='3 = fp32-class split-precision convs (parity mode), 1 = single fp16 pass')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel eagerly instead of replaying a CUDA graph')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    workload = f'FFHQ-{args.res} shgan_ffhq{args.res}_eval generator forward, batch {args.batch}/GPU, synthetic free-form masks, random-init weights'

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        cb, sec = cpu_reference_run(args.res, steps, max(1, min(args.warmup, 1)))
        line = dict(metric=METRIC, value=cb['value'], unit='images/s', n_gpus=args.gpus, steps=steps, warmup=1, ms_per_step=sec * 1e3,
                    higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                    config=dict(workload=workload, note='reference CPU path (PyTorch CPU conv) on the host cores; each step = 1 image'),
                    cpu_baseline=cb, e2e=dict(value=cb['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl shgan_b200 needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from shgan_b200 import _lib, kernels as K, synthetic as S

    dev = torch.device('cuda', local_rank)
    G = S.random_generator(args.res, seed=0, device=dev)
    eng = G.engine(passes=args.passes, impl=0)
    x_h, z_h = S.synthetic_batch(args.batch, args.res, seed=1000 + rank)
    x_pin, z_pin = x_h.pin_memory(), z_h.pin_memory()
    x_d, z_d = x_pin.to(dev), z_pin.to(dev)
    comp_pin = torch.empty((args.batch, 3, args.res, args.res), dtype=torch.uint8).pin_memory()

    # ---- instrumentation of the dominant kernel: CUDA events around every conv launch (current torch stream) ----
    conv_events, conv_flops = [], [0.0]
    record = [False]
    orig_conv = K.conv_igemm

    def conv_timed(srcs, w_hi, w_lo, taps, oh, ow, **kw):
        if not record[0]:
            return orig_conv(srcs, w_hi, w_lo, taps, oh, ow, **kw)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig_conv(srcs, w_hi, w_lo, taps, oh, ow, **kw)
        e.record()
        conv_events.append((s, e))
        n, _, _, c = srcs[0].shape
        conv_flops[0] += 2.0 * n * oh * ow * w_hi.shape[1] * c * len(taps)
    K.conv_igemm = conv_timed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return G.forward_composite(x_d, z_d, noise_mode='random')

    # end-to-end: what an eval loop does with this API -- pinned host batches in, uint8 composites out -- as a 2-deep
    # pipeline: the H2D copy of batch k+1 and the D2H copy of result k-1 run on a copy stream while batch k computes.
    copy_s = torch.cuda.Stream()
    dbuf = [(torch.empty_like(x_d), torch.empty_like(z_d)) for _ in range(2)]
    comp_dev = [torch.empty((args.batch, 3, args.res, args.res), dtype=torch.uint8, device=dev) for _ in range(2)]
    comp_pins = [comp_pin, torch.empty_like(comp_pin).pin_memory()]

    def run_e2e(k_steps):
        main = torch.cuda.current_stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        d2h = [torch.cuda.Event() for _ in range(2)]
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
            main.wait_event(ready[cur])
            if k >= 2:
                main.wait_event(d2h[cur])                      # comp_dev[cur] has been read back
            _, comp = G.forward_composite(dbuf[cur][0], dbuf[cur][1], noise_mode='random')
            comp_dev[cur].copy_(comp, non_blocking=True)       # the engine's output buffer is reused by the next replay
            freed[cur].record(main)
            done[cur].record(main)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done[cur])
                comp_pins[cur].copy_(comp_dev[cur], non_blocking=True)
                d2h[cur].record(copy_s)
        copy_s.synchronize()
        main.synchronize()

    # one eager (un-graphed) step counts the kernels of a step; the product path then replays them as a CUDA graph
    eng.graphs = False
    step_device()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    step_device()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0
    eng.graphs = not args.no_graphs
    for _ in range(max(args.warmup, 3)):
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

    # ---- same K steps again, launched eagerly with CUDA events around every conv launch (roofline of the dominant kernel;
    #      events cannot be read back from inside a replayed graph) ---------------------------------------------------
    eng.graphs = False
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
    conv_ms = sum(s.elapsed_time(e) for s, e in conv_events)
    n_conv = len(conv_events)
    eng.graphs = not args.no_graphs

    # ---- end-to-end timing (pinned host inputs, uint8 composite read back) ------------------------------------------
    run_e2e(3)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)                                       # ends with both streams synchronised: wall clock == device time
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if sampler else None

    tms = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])

    if rank == 0:
        peaks = _peaks()
        imgs = args.batch * world * args.steps
        value = imgs / (ms / 1e3)
        conv_tflops = conv_flops[0] / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'conv_tc_traffic.json')
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get('dram_bytes_per_launch_avg')
        line = dict(
            metric=METRIC, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f32' if args.passes == 3 else 'f16', data='synthetic',
            config=dict(workload=workload, resolution=args.res, batch_per_gpu=args.batch, global_batch=args.batch * world,
                        parallelism=f'dp{world} (batch-sharded, no collective in the forward)', noise_mode='random',
                        cuda_graph=not args.no_graphs,
                        precision='fp16 hi+lo split operands x3 tensor-core passes, fp32 accumulation (fp32-class, 1e-3 max-abs parity)'
                        if args.passes == 3 else 'single fp16 pass (NOT parity mode)',
                        l2_policy='working set per step (>1 GB of activations) exceeds the 126 MB L2; no explicit flush',
                        gflop_per_image=GFLOP_PER_IMAGE.get(args.res)),
            clocks=clocks,
            e2e=dict(value=imgs / (e2e_ms / 1e3), unit='images/s', ms_per_step=e2e_ms / args.steps,
                     h2d_bytes_per_step=int(x_pin.numel() * 4 + z_pin.numel() * 4), d2h_bytes_per_step=int(comp_pin.numel()),
                     api='model_zoo.comodgan.Generator.forward_composite(x, z); pinned host batches in, uint8 composites read back to pinned host; '
                         'H2D/D2H on a copy stream, 2-deep pipeline, all copies inside the timed region'),
            gpu_launches=int(launches),
            roofline=dict(bound='tensor', kernel='shgan_conv_igemm: conv_pair_kernel (tcgen05 cta_group::2, Co % 128 == 0) + conv_tc_kernel / conv_halo_kernel (single-CTA tcgen05), all 51 conv launches of a step',
                          achieved=conv_tflops, peak=peaks['tensor'], unit='TFLOP/s', frac=conv_tflops / peaks['tensor'],
                          traffic=traffic, peak_source=f'{peaks["src"]} bf16 sustained', launches_per_step=n_conv // max(args.steps, 1),
                          algorithmic_gflop_per_step=conv_flops[0] / max(args.steps, 1) / 1e9,
                          kernel_ms_per_step=conv_ms / args.steps, share_of_step=conv_ms / ms_instr,
                          instrumented_ms_per_step=ms_instr / args.steps,
                          executed_tensor_tflops=conv_tflops * (3 if args.passes == 3 else 1),
                          note='achieved = algorithmic conv FLOPs / summed CUDA-event durations of the conv launches over the same K steps launched eagerly '
                               '(value/ms_per_step come from the CUDA-graph replay of the identical launch sequence); '
                               'parity mode issues 3 fp16 MMA passes per algorithmic FLOP, so frac <= 1/3 by construction'),
            whole_step_algorithmic_tflops=GFLOP_PER_IMAGE.get(args.res, 0) * args.batch * world / (ms / args.steps) if args.res in GFLOP_PER_IMAGE else None,
        )
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_reference_run(args.res, 3, 1)
            line['cpu_baseline'] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
