"""CPU, world_size 2, gloo: the N>1 sampling path -- shard, sample independently, ONE all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagen_pytorch_b200.dist import sample_sharded


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeSampler:
    """Stands in for Imagen on the CPU box: 'images' are a deterministic function of the text embeds."""
    calls = 0

    def sample(self, text_embeds, text_masks=None, cond_scale=1.):
        _FakeSampler.calls += 1
        return text_embeds.mean(dim=(1, 2)).view(-1, 1, 1, 1).expand(-1, 3, 4, 4).contiguous() * cond_scale


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    te = torch.arange(n * 2 * 3, dtype=torch.float32).view(n, 2, 3)
    out = sample_sharded(_FakeSampler(), text_embeds=te, seed=0, cond_scale=2.)
    q.put((rank, out, _FakeSampler.calls))
    dist.barrier()
    dist.destroy_process_group()


def _run(n):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_sharded_sampling_gathers_global_batch():
    n = 6
    te = torch.arange(n * 2 * 3, dtype=torch.float32).view(n, 2, 3)
    expect = te.mean(dim=(1, 2)).view(-1, 1, 1, 1).expand(-1, 3, 4, 4) * 2.
    for rank, out, calls in _run(n):
        assert out.shape == (n, 3, 4, 4) and torch.equal(out, expect)
        assert calls == 1            # each rank sampled exactly once (its shard), no per-step communication


def test_ragged_shards_are_padded_and_trimmed():
    n = 5
    te = torch.arange(n * 2 * 3, dtype=torch.float32).view(n, 2, 3)
    expect = te.mean(dim=(1, 2)).view(-1, 1, 1, 1).expand(-1, 3, 4, 4) * 2.
    for rank, out, _ in _run(n):
        assert torch.equal(out, expect)


def test_fewer_samples_than_ranks_does_not_deadlock():
    """n < world: the trailing rank has an empty shard; it must still join the single all-gather."""
    n = 1
    te = torch.arange(n * 2 * 3, dtype=torch.float32).view(n, 2, 3)
    expect = te.mean(dim=(1, 2)).view(-1, 1, 1, 1).expand(-1, 3, 4, 4) * 2.
    res = _run(n)
    for rank, out, calls in res:
        assert out.shape == (1, 3, 4, 4) and torch.equal(out, expect)
    assert sorted(calls for _, _, calls in res) == [0, 1]    # only rank 0 sampled


def test_sample_in_chunks_splits_batched_arguments_like_the_trainer():
    from imagen_pytorch_b200.dist import sample_in_chunks

    class S:
        unconditional = False
        seen = []

        def sample(self, text_embeds=None, text_masks=None, cond_scale=1., return_all_unet_outputs=False):
            S.seen.append((text_embeds.shape[0], None if text_masks is None else text_masks.shape[0], cond_scale))
            img = text_embeds.mean(dim=(1, 2)).view(-1, 1, 1, 1).expand(-1, 3, 2, 2) * cond_scale
            return [img, img * 2] if return_all_unet_outputs else img

    te = torch.arange(7 * 2 * 3, dtype=torch.float32).view(7, 2, 3)
    tm = torch.ones(7, 2, dtype=torch.bool)
    full = S().sample(text_embeds=te, text_masks=tm, cond_scale=3.)
    S.seen.clear()
    out = sample_in_chunks(S(), text_embeds=te, text_masks=tm, cond_scale=3., max_batch_size=3)
    assert torch.equal(out, full) and S.seen == [(3, 3, 3.), (3, 3, 3.), (1, 1, 3.)]
    outs = sample_in_chunks(S(), text_embeds=te, cond_scale=1., return_all_unet_outputs=True, max_batch_size=4)
    assert len(outs) == 2 and outs[0].shape[0] == 7 and torch.equal(outs[1], outs[0] * 2)
    assert torch.equal(sample_in_chunks(S(), text_embeds=te, cond_scale=3.), full)     # no max_batch_size: one call
