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
