"""Data-parallel sampling: independent sample batches sharded across the GPUs of one box.

The reference has no data-parallel sampling (ImagenTrainer.sample runs the full batch on every rank,
trainer.py:947-961).  Every op of the path is per-sample (RMS/LayerNorm per
pixel/token, per-sample quantile, per-sample CFG pair), so the global batch is split into contiguous
shards, each rank runs its own CUDA-graphed t-loop with ZERO communication, and the finished images
are collected with ONE all-gather (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced shard [lo, hi) of n items for `rank` of `world` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_images(local, counts, group=None):
    """Single all-gather of the finished images; ragged shards are padded to the largest one."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat((local, local.new_zeros((mx - local.shape[0], *local.shape[1:]))), dim=0)
    out = torch.empty((world * mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * mx:r * mx + c] for r, c in enumerate(counts)], dim=0)


def sample_sharded(sampler, *, text_embeds, text_masks=None, seed=None, gather=True, group=None, sample_fn=None, **sample_kwargs):
    """Shard `text_embeds` (global batch first dim) over the ranks of `group`, sample locally, all-gather.

    RNG: rank r seeds its generator with seed + r (documented: results differ from a 1-GPU run of the
    global batch, which would draw one noise tensor for all samples).  sample_fn(sampler, **kw) lets tests
    substitute the sampling call (the CPU/gloo tests cover the sharding + gather logic only)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = text_embeds.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    counts = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    if seed is not None:
        torch.manual_seed(seed + rank)
    kw = dict(sample_kwargs)
    kw['text_embeds'] = text_embeds[lo:hi]
    if text_masks is not None:
        kw['text_masks'] = text_masks[lo:hi]
    fn = sample_fn if sample_fn is not None else (lambda s, **k: s.sample(**k))
    local = fn(sampler, **kw)
    if world == 1 or not gather:
        return local
    return all_gather_images(local, counts, group=group)
