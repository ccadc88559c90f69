"""Two-rank NCCL test (`-m gpu`, needs 2 GPUs; skipped on a 1-GPU box) of the eval loop's only exchange: `parallel.EvalLoop`
with the real generator on each rank's own GPU and ONE all_gather of the detector features at the end
(replaces the per-batch broadcasts of lib/evaluator/eva_base.py:96-194, eva_fid.py:213-223,253-259).

  * batch size 1: every item is its own batch, so the gathered features must be BIT-identical to a single-process run over
    the whole dataset (per-item latents, `noise_mode='const'`), in dataset order with the wrap-around padding dropped;
  * batch size 3 with an item count that does not divide: each rank's own rows come back unchanged at their dataset
    positions, both ranks hold the same matrix (the order check inside EvalLoop.run runs on the device).
The detector is a device-side stand-in (the Inception TorchScript is not available offline)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
RES, N_ITEMS = 128, 11


def _detector(u8):
    """[b,3,R,R] uint8 -> [b,48] features: 4x4 average pooling of each channel (stand-in for the Inception pool3 features)."""
    return torch.nn.functional.adaptive_avg_pool2d(u8.float(), 4).flatten(1)


def _dataset(i):
    g = torch.Generator().manual_seed(1000 + i)
    real = torch.rand(3, RES, RES, generator=g) * 2 - 1
    mask = (torch.rand(RES // 8, RES // 8, generator=g) > 0.4).float().repeat_interleave(8, 0).repeat_interleave(8, 1)
    return real, mask


def _build(device):
    from oracle import shgan_oracle as O          # checker-side helper: synthetic weights only
    import helpers as H
    sd = O.synthetic_state_dict(RES, seed=11, ch_base=8192, ch_max=64)
    return H.build_generator(RES, sd, 8192, 64, device=device)


def _run(device, batch_size):
    from shgan_b200 import parallel as PL
    loop = PL.EvalLoop(_build(device), _detector, batch_size=batch_size, device=device, z_seed=3)
    return loop, loop.run(_dataset, N_ITEMS, noise_mode='const')


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    device = torch.device('cuda', rank)           # as the reference eval: no torch.cuda.set_device (shgan_default.py:164)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=device)
    from shgan_b200 import parallel as PL
    out = {}
    for bs in (1, 3):
        loop, (fake, real) = _run(device, bs)
        assert fake.device == device and fake.dtype == torch.float64 and fake.shape == (N_ITEMS, 48)
        out[bs] = (fake.cpu(), real.cpu(), loop.last_gather_bytes)
    # batch size 3: my own rows, recomputed locally batch by batch, sit at their dataset positions
    mine = PL.shard_indices(N_ITEMS, rank, world)
    q.put((rank, mine, out))
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_eval_loop_world2_nccl_feature_gather():
    world, port = 2, 29641
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # both ranks hold the same gathered matrices
    for bs in (1, 3):
        for k in range(2):
            assert torch.equal(res[0][2][bs][k], res[1][2][bs][k])
        assert res[0][2][bs][2] == world * 6 * (2 * 48 + 1) * 8          # ceil(11 / 2) rows per rank, one gather
    # single-process run over the whole dataset on this process's GPU
    _, (fake1, real1) = _run(torch.device('cuda', 0), 1)
    assert torch.equal(fake1.cpu(), res[0][2][1][0]) and torch.equal(real1.cpu(), res[0][2][1][1])
    # real-image features do not depend on the batching at all
    assert torch.equal(real1.cpu(), res[0][2][3][1])
    # batch size 3: close to the batch-1 result (the batch-global style normaliser, a reference quirk the engine keeps --
    # stylegan.py:145-155 -- makes the images depend on the batch composition), never identical to a stale buffer
    d = (res[0][2][3][0] - fake1.cpu()).abs().max()
    assert torch.isfinite(res[0][2][3][0]).all() and float(d) < 64.0
