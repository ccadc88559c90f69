"""world_size-2 gloo tests (CPU) of the eval sharding / feature gather host logic (shgan_b200/parallel.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _reference_order(per_rank_batches):
    """The reference's per-batch broadcast + zipzap_arrange concatenation (eva_base.py:196-225), on python lists."""
    out = []
    nb = max(len(b) for b in per_rank_batches)
    for k in range(nb):
        chunk = [b[k] for b in per_rank_batches if k < len(b)]
        maxlen = max(len(c) for c in chunk)
        tot = sum(len(c) for c in chunk)
        cnt = 0
        for i in range(maxlen):
            for c in chunk:
                if i < len(c) and cnt < tot:
                    out.append(c[i])
                    cnt += 1
    return out


def _worker(rank, world, port, n_items, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from shgan_b200 import parallel as PL
    mine = PL.shard_indices(n_items, rank, world)
    feats = torch.tensor([[float(i), float(i) * 2 + 1] for i in mine], dtype=torch.float64)
    full = PL.gather_features(feats, n_items)
    q.put((rank, mine, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [10, 11, 7])
def test_shard_and_gather_world2(n_items):
    world, port = 2, 29500 + n_items
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = np.array([[float(i), float(i) * 2 + 1] for i in range(n_items)])
    for rank, mine, full in res:
        assert mine[:len(range(rank, n_items, world))] == list(range(rank, n_items, world))
        assert full.shape == (n_items, 2) and np.array_equal(full, exp)
    # same order as the reference's per-batch broadcast + zipzap with batch size 2
    bs = 2
    per_rank = [[m[i:i + bs] for i in range(0, len(m), bs)] for _, m, _ in res]
    assert _reference_order(per_rank)[:n_items] == list(range(n_items))


def test_shard_indices_matches_reference_rule():
    from shgan_b200.parallel import shard_indices
    for n, w in [(10, 4), (8, 4), (5, 8), (36500, 8)]:
        allr = [shard_indices(n, r, w) for r in range(w)]
        per = -(-n // w)
        assert all(len(a) == per for a in allr)
        flat = [allr[i % w][i // w] for i in range(per * w)]
        assert flat[:n] == list(range(n)) and flat[n:] == list(range(per * w - n))     # wrap-around padding


def test_fid_from_features():
    from shgan_b200.parallel import fid_from_features
    g = np.random.default_rng(0)
    a = g.standard_normal((400, 16))
    assert abs(fid_from_features(a, a)) < 1e-6
    b = a + 0.5
    assert abs(fid_from_features(a, b) - 16 * 0.25) < 1e-6


class _FakeG:
    """Stands in for the generator in the host-logic test of EvalLoop: the 'composite' is a function of (x, z) only."""
    z_dim = 8

    def forward_composite(self, x, z, noise_mode='random'):
        v = (x[:, 1:4].mean(dim=(1, 2, 3)) * 40 + z.sum(dim=1)).reshape(-1, 1, 1, 1)
        img = v.expand(-1, 3, 4, 4).contiguous()
        return img, (img * 8 + 128).clamp(0, 255).to(torch.uint8)


def _dataset(i):
    g = torch.Generator().manual_seed(500 + i)
    return torch.rand(3, 4, 4, generator=g) * 2 - 1, (torch.rand(4, 4, generator=g) > 0.3).float()


def _eval_worker(rank, world, port, n_items, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from shgan_b200 import parallel as PL
    loop = PL.EvalLoop(_FakeG(), lambda u8: u8.float().mean(dim=(2, 3)), batch_size=2, device='cpu', z_seed=7)
    fake, real = loop.run(_dataset, n_items)
    q.put((rank, fake.numpy(), real.numpy(), loop.last_gather_bytes))
    dist.barrier()
    dist.destroy_process_group()


def test_eval_loop_world2_matches_world1():
    """EvalLoop over 2 gloo ranks == the same loop on one rank: per-item latents do not depend on the sharding, the single
    gather restores dataset order (its index column is checked inside run()) and drops the wrap-around duplicate."""
    from shgan_b200 import parallel as PL
    n_items = 7
    one = PL.EvalLoop(_FakeG(), lambda u8: u8.float().mean(dim=(2, 3)), batch_size=2, device='cpu', z_seed=7)
    f1, r1 = one.run(_dataset, n_items)
    world, port = 2, 29533
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fake, real, nbytes in res:
        assert fake.shape == (n_items, 3) and np.array_equal(fake, f1.numpy()) and np.array_equal(real, r1.numpy())
        assert nbytes == world * 4 * (2 * 3 + 1) * 8
    # different items get different latents, whatever the rank that evaluates them
    assert len({float(one.latent(i)[0]) for i in range(n_items)}) == n_items
