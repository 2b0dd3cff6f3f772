"""CPU: host-side schedule tables, quantile rank arithmetic, weight packing and shard logic."""
import math

import pytest
import torch

from imagen_pytorch_b200 import ElucidatedImagen, Imagen, Unet
from imagen_pytorch_b200.dist import shard_bounds
from imagen_pytorch_b200.imagen import GaussianDiffusionContinuousTimes, quantile_ranks
from imagen_pytorch_b200 import ops
from oracle import sampler_ref


@pytest.mark.parametrize('n', [12288, 3072, 196608, 10])
def test_quantile_ranks_reproduce_torch_quantile(n):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(3, n, generator=g).abs()
    lo, hi, w = quantile_ranks(n, 0.95, 'cpu')
    s = x.sort(dim=-1).values
    a, b = s[:, lo], s[:, hi]
    mine = torch.where(torch.tensor(w) < 0.5, a + w * (b - a), b - (b - a) * (1 - w))
    assert torch.equal(mine, torch.quantile(x, 0.95, dim=-1))


@pytest.mark.parametrize('schedule', ['cosine', 'linear'])
def test_ddpm_coefficient_table_matches_oracle_posterior(schedule):
    T = 7
    sched = GaussianDiffusionContinuousTimes(noise_schedule=schedule, timesteps=T)
    coefs, log_snr = sched.ddpm_coefficients('cpu')
    fn = sampler_ref.LOG_SNR[schedule]
    x_t, x0 = torch.randn(1, 3, 4, 4), torch.randn(1, 3, 4, 4)
    for i, (t, tn) in enumerate(sampler_ref.sampling_timesteps(T, 1)):
        mean, _, log_var = sampler_ref.q_posterior(fn, x0, x_t, t, tn)
        sigma, alpha, _, alpha_next, c, noise_std = coefs[i, :6]
        mine = alpha_next * (x_t * (1 - c) / alpha + c * x0)
        assert torch.equal(mine, mean)
        assert torch.equal(log_snr[i], fn(t)[0])
        expect_std = (0.5 * log_var).exp().flatten()[0] * (0. if tn.item() == 0 else 1.)
        assert torch.equal(noise_std, expect_std)
    assert coefs[-1, 5] == 0            # no noise on the last step


@pytest.mark.parametrize('schedule', ['cosine', 'linear'])
def test_repaint_coefficient_table_matches_oracle_q_sample(schedule):
    """(alpha_t, sigma_t) of q_sample and (c1, c2, alpha_from) of q_sample_from_to, as b200_inpaint_mix / b200_renoise consume them:
    x*c1 + (noise*c2)/alpha_from must reproduce the oracle's q_sample_from_to(x, t_next -> t) bit for bit."""
    T = 6
    sched = GaussianDiffusionContinuousTimes(noise_schedule=schedule, timesteps=T)
    rp = sched.repaint_coefficients('cpu')
    fn = sampler_ref.LOG_SNR[schedule]
    g = torch.Generator().manual_seed(3)
    x, noise = torch.randn(2, 3, 4, 4, generator=g), torch.randn(2, 3, 4, 4, generator=g)
    steps = list(sampler_ref.sampling_timesteps(T, 2))
    for i, (t, t_next) in enumerate(steps):
        a, s_, c1, c2, af = rp[i]
        assert torch.equal(a * x + s_ * noise, sampler_ref.q_sample(fn, x, t, noise))
        if i < T - 1:                                            # the last step (t_next = 0) is never re-noised (:2271)
            assert torch.equal(x * c1 + (noise * c2) / af, sampler_ref.q_sample_from_to(fn, x, t_next, t, noise))


def test_edm_tables_match_oracle_schedule():
    u = Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64)
    el = ElucidatedImagen(u, image_sizes=16, text_embed_dim=64, num_sample_steps=5)
    hp = el.hparams[0]
    coefs, times, init_sigma = el._edm_tables(hp, hp.sigma_min, hp.sigma_max, 'cpu')
    sig = sampler_ref.edm_sample_schedule(5, 7, 0.002, 80)
    assert init_sigma == sig[0] and coefs.shape == (5, 16) and times.numel() == 9   # 5 + 4 network evaluations
    gamma = min(80 / 5, math.sqrt(2) - 1)
    s0 = sig[0].item()
    assert abs(coefs[0, 2].item() - (s0 + gamma * s0 if 0.05 <= s0 <= 50 else s0)) < 1e-5
    assert coefs[-1, 12] == 0 and coefs[0, 12] == 1                                  # has_second
    assert torch.allclose(times[0], torch.log(coefs[0, 2]) * 0.25)


def test_weight_packing_layout():
    W = torch.arange(2 * 5 * 3 * 3, dtype=torch.float32).view(2, 5, 3, 3)
    segs, mats = ops.conv_segments(W, [2, 3])
    assert len(segs) == 18 and segs[0] == (0, -1, -1) and segs[1] == (1, -1, -1) and segs[-1] == (1, 1, 1)
    assert torch.equal(mats[0], W[:, :2, 0, 0]) and torch.equal(mats[-1], W[:, 2:, 2, 2])


@pytest.mark.parametrize('n,world', [(512, 8), (10, 4), (3, 8), (16, 1)])
def test_shard_bounds_partition_the_batch(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_sample_argument_validation_mirrors_reference():
    u = Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64)
    im = Imagen(u, image_sizes=16, text_embed_dim=64, timesteps=2, cond_drop_prob=0.)
    with pytest.raises(AssertionError):
        im.sample(text_embeds=torch.randn(1, 8, 32), use_tqdm=False)          # wrong embedding dim
    with pytest.raises(AssertionError):
        im.sample(use_tqdm=False)                                               # text required
    with pytest.raises(AssertionError):
        im.sample(texts=['a cat', ''], use_tqdm=False)                          # 'text cannot be empty' (:2327)
    with pytest.raises(AssertionError):                                           # cond_images for a U-Net built without cond_images_channels (:1555)
        im.sample(text_embeds=torch.randn(1, 8, 64), cond_images=torch.zeros(1, 3, 16, 16), use_tqdm=False)
    with pytest.raises(NotImplementedError):
        im.sample(text_embeds=torch.randn(1, 8, 64), video_frames=4, use_tqdm=False)                                              # video: out of the hot path
    with pytest.raises(AssertionError):                                           # inpainting batch must match the text batch (:2343-2350)
        im.sample(text_embeds=torch.randn(1, 8, 64), inpaint_images=torch.zeros(2, 3, 16, 16), inpaint_masks=torch.zeros(2, 16, 16), use_tqdm=False)


def test_plan_compiles_on_the_meta_device_without_kernels():
    """The launch-plan compiler (weight packing through the staged host arena, activation arena, launch list) is pure host
    logic: build it on torch's meta device and check the bookkeeping."""
    import imagen_pytorch_b200 as b2
    from imagen_pytorch_b200.unet import UnetPlan
    u = b2.Unet(dim=32, dim_mults=(1, 2, 4, 8), text_embed_dim=64, max_text_len=24)
    plan = UnetPlan(u, 4, 2, 32, 32, 6, torch.device('meta'))
    assert plan.n_launches > 100 and len(plan._ops) > 100
    assert plan.fingerprint == u._fingerprint()
    assert plan.n_ctx == 2 + 32 + 4                           # time tokens + perceiver latents (32 + 4 mean-pooled)
    # every weight operand lives in the staged arena, every activation in the zero-filled one: a handful of chunks in total
    assert len(plan._wts.chunks) <= 4 and len(plan._act.chunks) <= 4
    with torch.no_grad():
        next(u.parameters()).add_(1.0)                        # in-place update -> fingerprint changes -> plans are rebuilt
    assert plan.fingerprint != u._fingerprint()


def test_unet_plan_cache_survives_noop_module_moves():
    import imagen_pytorch_b200 as b2
    u = b2.Unet(dim=32, dim_mults=(1, 2))
    fp = u._fingerprint()
    u._plans['sentinel'] = type('P', (), {'fingerprint': fp})()
    u.to('cpu')                                               # Imagen.sample() calls unets.to(device) on every call
    assert 'sentinel' in u._plans and u._fingerprint() == fp


class _ToyTokenizer:
    """Whitespace tokenizer with the HF batch_encode_plus contract (right padding, 'longest')."""

    def batch_encode_plus(self, texts, return_tensors='pt', padding='longest', max_length=256, truncation=True):
        ids = [[(hash(w) % 97) + 3 for w in t.split()][:max_length - 1] + [1] for t in texts]      # 1 = </s>
        n = max(len(i) for i in ids)
        out = type('Enc', (), {})()
        out.input_ids = torch.tensor([i + [0] * (n - len(i)) for i in ids])
        out.attention_mask = torch.tensor([[1] * len(i) + [0] * (n - len(i)) for i in ids])
        return out


def _toy_t5(dim=64):
    from transformers import T5Config, T5EncoderModel
    torch.manual_seed(0)
    return T5EncoderModel(T5Config(vocab_size=100, d_model=dim, d_kv=16, d_ff=128, num_layers=2, num_heads=4)).eval()


def test_t5_mirror_masks_padding_and_cache_reproduces_batches():
    """imagen_pytorch_b200.t5 (mirror of the reference's t5.py): padding embeddings are exactly zero, the mask is the tokenizer's,
    and the per-prompt LRU returns the same embeddings for a prompt whatever batch it was first encoded in."""
    from imagen_pytorch_b200 import t5
    t5.register_text_encoder('toy-t5', _toy_t5(), _ToyTokenizer())
    assert t5.get_encoded_dim('toy-t5') == 64 and t5.get_encoded_dim('google/t5-v1_1-xl') == 2048
    texts = ['a photo of a cat', 'dog', 'a very long prompt about nothing in particular']
    emb, mask = t5.t5_encode_text(texts, name='toy-t5', return_attn_mask=True)
    emb, mask = emb.cpu(), mask.cpu()
    assert emb.shape == (3, 9, 64) and mask.dtype == torch.bool and mask.sum(1).tolist() == [6, 2, 9]
    assert emb[~mask].abs().max() == 0 and emb[mask].abs().min() > 0
    cache = t5.TextEmbedCache(capacity=8)
    e1, m1 = cache.encode(texts, name='toy-t5')
    assert (cache.hits, cache.misses) == (0, 3)
    assert torch.equal(m1, mask) and torch.allclose(e1, emb, atol=1e-5)
    e2, m2 = cache.encode(['dog', 'a photo of a cat'], name='toy-t5')           # all hits; different batch => different padded length
    assert (cache.hits, cache.misses) == (2, 3) and e2.shape == (2, 6, 64)
    ref, refm = t5.t5_encode_text(['dog', 'a photo of a cat'], name='toy-t5', return_attn_mask=True)
    assert torch.equal(m2, refm.cpu()) and torch.allclose(e2, ref.cpu(), atol=1e-5)


def test_sample_with_texts_goes_through_encode_text_hook():
    """Imagen.sample(texts=...) encodes through self.encode_text (imagen_pytorch.py:2326-2332) -- here it reaches the device check
    with text_embeds of the encoder's width; with the cache enabled the second call is served from the LRU."""
    from imagen_pytorch_b200 import t5, B200Error
    t5.register_text_encoder('toy-t5', _toy_t5(), _ToyTokenizer())
    im = Imagen(Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64), image_sizes=16, text_encoder_name='toy-t5', timesteps=2)
    assert im.text_embed_dim == 64
    cache = im.enable_text_embed_cache()
    for _ in range(2):
        with pytest.raises(B200Error):                                          # CPU box: the sampler refuses to run, after the text was encoded
            im.sample(texts=['a cat', 'a dog on a skateboard'], use_tqdm=False)
    assert (cache.hits, cache.misses) == (2, 2)


def test_trained_sampler_loads_trainer_checkpoint_and_swaps_ema_unets(tmp_path):
    """ImagenTrainer.save layout (trainer.py:677-736) -> TrainedSampler.load: online weights into imagen.unets, 'i.ema_model.*' into the EMA
    twins; sample() runs on the EMA U-Nets unless use_non_ema (trainer.py:846-869, :947-961) and chunks by max_batch_size (:188-206)."""
    from imagen_pytorch_b200 import TrainedSampler
    kw = dict(dim=32, dim_mults=(1, 2), text_embed_dim=64)
    im = Imagen((Unet(**kw), Unet(**kw)), image_sizes=(16, 32), text_embed_dim=64, timesteps=2)
    sds = [{k: torch.randn_like(v) for k, v in u.state_dict().items()} for u in im.unets]
    emas = [{k: torch.randn_like(v) for k, v in u.state_dict().items()} for u in im.unets]
    ckpt = dict(model={f'unets.{i}.{k}': v for i, sd in enumerate(sds) for k, v in sd.items()},
                ema={**{f'{i}.ema_model.{k}': v for i, sd in enumerate(emas) for k, v in sd.items()},
                     **{f'{i}.online_model.{k}': v for i, sd in enumerate(sds) for k, v in sd.items()},     # older ema_pytorch versions save these too
                     '0.initted': torch.tensor([True]), '0.step': torch.tensor([7]), '1.initted': torch.tensor([True]), '1.step': torch.tensor([7])},
                version='1.26.2', steps=torch.tensor([7, 7]), optim0={'state': {}}, scaler0={})
    path = tmp_path / 'checkpoint.pt'
    torch.save(ckpt, path)
    ts = TrainedSampler(im)
    steps, version = ts.load(str(path))
    assert version == '1.26.2' and steps.tolist() == [7, 7]
    for i in range(2):
        assert all(torch.equal(im.unets[i].state_dict()[k], sds[i][k]) for k in sds[i])
        assert all(torch.equal(ts.ema_unets[i].state_dict()[k], emas[i][k]) for k in emas[i])
    seen = []
    im.sample = lambda **k: (seen.append((im.unets is ts.ema_unets, k['text_embeds'].shape[0])), torch.zeros(k['text_embeds'].shape[0], 3, 4, 4))[1]
    out = ts.sample(text_embeds=torch.randn(5, 8, 64), cond_scale=2., max_batch_size=2)
    assert out.shape[0] == 5 and seen == [(True, 2), (True, 2), (True, 1)] and im.unets is not ts.ema_unets     # swapped in, chunked, restored
    seen.clear()
    ts.sample(text_embeds=torch.randn(3, 8, 64), use_non_ema=True)
    assert seen == [(False, 3)]
    with pytest.raises(RuntimeError):
        TrainedSampler(im).load(dict(model=ckpt['model'], version='x'))                 # no EMA weights in the checkpoint
