"""Batch-sharded evaluation of the generator on the GPUs of one node ("next" row N1 of SURVEY.md section 8f).

The generator forward needs no communication: weights are replicated, images are independent.  The only exchange of
the eval loop is the Inception-feature gather, done here as ONE all-gather at the end of the run instead of the
reference's per-batch, per-rank, per-array triple broadcasts (lib/evaluator/eva_base.py:96-194, eva_fid.py:213-223).

    shard_indices    <- DistributedSampler(shuffle=False, extend=True), lib/data_factory/common/ds_sampler.py:58-68
    gather_features  <- base_evaluator.sync + zipzap_arrange + truncation to sample_n, eva_base.py:96-225, eva_fid.py:253-259
    fid_from_features<- fid_evaluator.compute_fid, eva_fid.py:252-277
    EvalLoop         <- eval_stage.__call__'s batch loop, lib/experiments/shgan_default.py:257-295

Works with any torch.distributed backend (nccl on the GPUs; gloo in the CPU tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world, extend=True):
    """Indices of the dataset items that `rank` evaluates: r, r+W, r+2W, ...; with extend=True the list is padded by
    wrapping around to the front so that every rank gets ceil(n/W) items."""
    per = n_items // world
    if extend and per * world != n_items:
        per += 1
    total = per * world
    idx = list(range(n_items))
    idx = idx + idx[:total - len(idx)] if extend else idx[:total]
    return idx[rank:len(idx):world]


def gather_features(local, n_items, group=None):
    """local: [n_local, D] features of this rank's shard_indices() items, in shard order.  Returns the [n_items, D]
    matrix in dataset order on every rank (one all_gather; wrap-around duplicates dropped)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n_items]
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    # item i of rank r is dataset item r + i*W: interleave the ranks ("zipzap" order of the reference)
    stacked = torch.stack(parts, dim=1)                       # [n_local, W, D]
    return stacked.reshape(-1, local.shape[1])[:n_items]


def _sqrtm_psd_product(a, b):
    """trace-compatible matrix square root of a @ b for covariance matrices (scipy.linalg.sqrtm when available)."""
    try:
        import scipy.linalg
        s = scipy.linalg.sqrtm(a @ b)     # newer SciPy dropped the (sqrtm, err) tuple of `disp=False`
        return np.real(s[0] if isinstance(s, tuple) else s)
    except ImportError:  # pragma: no cover
        w, v = np.linalg.eig(a @ b)
        return np.real((v * np.sqrt(w.astype(complex))) @ np.linalg.inv(v))


def fid_from_features(fake, real):
    """Frechet distance between two feature sets, float64 on the host (rank-0 work in the reference)."""
    fake = np.asarray(fake, np.float64)
    real = np.asarray(real, np.float64)
    mu_f, mu_r = fake.mean(0), real.mean(0)
    sig_f = fake.T @ fake / fake.shape[0] - np.outer(mu_f, mu_f)
    sig_r = real.T @ real / real.shape[0] - np.outer(mu_r, mu_r)
    s = _sqrtm_psd_product(sig_f, sig_r)
    return float(np.square(mu_f - mu_r).sum() + np.trace(sig_f + sig_r - 2 * s))


class EvalLoop:
    """Runs the generator over this rank's shard and accumulates detector features ON THE DEVICE.

    dataset(i) -> (real [3,R,R] in [-1,1], mask [R,R] in {0,1}); detector(uint8 images [b,3,R,R]) -> [b,D] features.
    The images never leave the GPU: composite + uint8 quantisation are fused into the last generator kernel."""

    def __init__(self, G, detector, batch_size, device, z_seed=0):
        self.G, self.detector, self.batch_size, self.device = G, detector, batch_size, device
        self.z_seed = int(z_seed)

    def latent(self, index):
        """z of dataset item `index`: a function of (z_seed, index) only, so every item gets its own latent whatever the
        rank count and the sharding (the reference seeds per rank, rnd_seed*gpu_count+RANK, lib/experiments/
        shgan_default.py:160-166: independent across ranks, but world-size dependent)."""
        g = torch.Generator(device='cpu').manual_seed((self.z_seed << 32) + int(index))
        return torch.randn([self.G.z_dim], generator=g)

    def prepare(self, real, mask):
        """x = cat([mask - 0.5, real * mask]) (shgan_default.py:269-274): one kernel on the GPU; plain torch ops only in the
        CPU tests of the host logic."""
        if real.is_cuda:
            from . import kernels as K
            return K.prepare_input(real.contiguous().float(), mask.contiguous().float())
        return torch.cat([mask - 0.5, real * mask], dim=1)

    def run(self, dataset, n_items, noise_mode='random'):
        """-> (fake [n_items, D], real [n_items, D]) float64 detector features in dataset order, on every rank.
        ONE all_gather for the whole run: fake features, real features and the item index travel as one
        [n_local, 2D+1] matrix; the gathered index column must come back as 0..n_items-1 (checked on the device)."""
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        mine = shard_indices(n_items, rank, world)
        rows = []
        for b0 in range(0, len(mine), self.batch_size):
            idx = mine[b0:b0 + self.batch_size]
            items = [dataset(i) for i in idx]
            real = torch.stack([it[0] for it in items]).to(self.device, non_blocking=True)
            mask = torch.stack([it[1] for it in items])[:, None].to(self.device, non_blocking=True)
            z = torch.stack([self.latent(i) for i in idx]).to(self.device, non_blocking=True)
            if real.is_cuda and hasattr(self.G, 'forward_inpaint'):
                # mask / erase / concat (shgan_default.py:269-274) fused into the first kernel of the generator
                _, fake_u8 = self.G.forward_inpaint(real.contiguous().float(), mask.contiguous().float(), z, noise_mode=noise_mode)
            else:
                _, fake_u8 = self.G.forward_composite(self.prepare(real, mask), z, noise_mode=noise_mode)
            f_fake = self.detector(fake_u8).to(torch.float64)
            f_real = self.detector((real * 127.5 + 127.5).clamp(0, 255).to(torch.uint8)).to(torch.float64)
            col = torch.tensor(idx, dtype=torch.float64, device=f_fake.device)[:, None]
            rows.append(torch.cat([f_fake, f_real, col], dim=1))
        full = gather_features(torch.cat(rows), n_items)
        self.last_gather_bytes = int(world * len(mine) * full.shape[1] * 8)
        d = (full.shape[1] - 1) // 2
        order_ok = torch.equal(full[:, -1], torch.arange(n_items, dtype=torch.float64, device=full.device))
        if not order_ok:
            raise RuntimeError('feature gather returned items out of dataset order')
        return full[:, :d], full[:, d:2 * d]
