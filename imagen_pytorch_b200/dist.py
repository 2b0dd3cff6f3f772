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
    local = fn(sampler, **kw) if hi > lo else None           # fewer samples than ranks: the trailing ranks have nothing to sample
    if world == 1 or not gather:
        return local if local is not None else text_embeds.new_zeros((0,))
    if n < world:
        # every rank knows (n < world) => some shards are empty: rank 0 (never empty for n >= 1) publishes the image shape so the
        # empty ranks can contribute a zero-row tensor to the single all-gather instead of blocking it
        assert n >= 1, 'sample_sharded needs at least one sample'
        dev = local.device if local is not None else getattr(sampler, 'device', text_embeds.device)
        shp = torch.tensor(list(local.shape[1:]) if rank == 0 else [0, 0, 0], dtype=torch.int64, device=dev)
        dist.broadcast(shp, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if local is None:
            local = torch.zeros((0, *shp.tolist()), dtype=torch.float32, device=dev)
    return all_gather_images(local, counts, group=group)


def num_to_groups(num, divisor):
    groups, remainder = divmod(num, divisor)
    return [divisor] * groups + ([remainder] if remainder > 0 else [])


def sample_in_chunks(sampler, *args, max_batch_size=None, **kwargs):
    """`ImagenTrainer.sample(..., max_batch_size=)` (trainer.py:188-206, :947-961) for the B200 samplers: split every batched
    argument (tensors and lists, first dim / len == batch) into chunks of at most `max_batch_size`, sample each chunk, and
    concatenate -- per U-Net stage when `return_all_unet_outputs=True`.  Non-batched arguments are passed to every chunk.
    Chunks of equal size reuse the same UnetPlan and captured step graph; a ragged last chunk compiles one more plan."""
    if max_batch_size is None:
        return sampler.sample(*args, **kwargs)
    if getattr(sampler, 'unconditional', False):
        sizes = num_to_groups(kwargs.get('batch_size', 1), max_batch_size)
        outputs = [sampler.sample(*args, **{**kwargs, 'batch_size': b}) for b in sizes]
    else:
        vals = (*args, *kwargs.values())
        first = next((v for v in vals if isinstance(v, torch.Tensor)), None)
        assert first is not None, 'sample_in_chunks needs at least one batched tensor argument (e.g. text_embeds)'
        n = len(first)
        bounds = [(i, min(i + max_batch_size, n)) for i in range(0, n, max_batch_size)]

        def cut(v, lo, hi):
            if isinstance(v, torch.Tensor) and v.ndim > 0 and v.shape[0] == n:
                return v[lo:hi]
            if isinstance(v, (list, tuple)) and len(v) == n and not isinstance(v, str):
                return v[lo:hi]
            return v

        outputs = [sampler.sample(*[cut(a, lo, hi) for a in args], **{k: cut(v, lo, hi) for k, v in kwargs.items()}) for lo, hi in bounds]
    if isinstance(outputs[0], torch.Tensor):
        return torch.cat(outputs, dim=0)
    return [torch.cat(t, dim=0) for t in zip(*outputs)]
