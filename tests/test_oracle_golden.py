"""CPU: the oracle (oracle/) reproduces the golden vectors generated from the live reference
(oracle/make_golden.py).  These fixtures are the reference's outputs, not the oracle's."""
import torch

from oracle import unet_ref, sampler_ref
from tests.helpers import load_golden, synth_weights

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def test_unet_base_forward_matches_reference_golden():
    g = load_golden('unet_base_dim32.pt')
    sd = synth_weights('test_base', g['wseed'])
    cfg = unet_ref.unet_config(**g['kwargs'])
    kw = dict(text_embeds=g['text_embeds'], text_mask=g['text_mask'])
    with torch.no_grad():
        out = unet_ref.unet_forward(sd, cfg, g['x'], g['t'], **kw)
        out_null = unet_ref.unet_forward(sd, cfg, g['x'], g['t'], cond_drop_prob=1., **kw)
    assert out.abs().max() > 1.0                       # not the vacuous zero-init case
    assert (out - g['out_cond']).abs().max() < 1e-4
    assert (out_null - g['out_null']).abs().max() < 1e-4
    assert (out - out_null).abs().max() > 0.1          # text conditioning actually matters


def test_unet_sr_forward_matches_reference_golden():
    g = load_golden('unet_sr_dim32.pt')
    sd = synth_weights('test_sr', g['wseed'])
    cfg = unet_ref.unet_config(**g['kwargs'], lowres_cond=True)
    with torch.no_grad():
        out = unet_ref.unet_forward(sd, cfg, g['x'], g['t'], text_embeds=g['text_embeds'], text_mask=g['text_mask'],
                                    lowres_cond_img=g['lowres_cond_img'], lowres_noise_times=g['lowres_noise_times'])
    assert (out - g['out']).abs().max() < 1e-4


def test_ddpm_sample_matches_reference_golden():
    g = load_golden('ddpm_sample_dim32.pt')
    sd = synth_weights('test_base', g['wseed'])
    cfg = unet_ref.unet_config(**g['kwargs'])
    torch.manual_seed(g['seed'])
    trace = []
    with torch.no_grad():
        out = sampler_ref.imagen_sample([(sd, cfg)], (32,), text_embeds=g['text_embeds'], timesteps=g['timesteps'],
                                        cond_scale=g['cond_scale'], trace=trace)
    assert (out - g['out']).abs().max() < 1e-4
    assert (torch.stack(trace) - g['trace']).abs().max() < 1e-3
    assert 0. <= out.min() and out.max() <= 1.


def test_ddpm_cascade_matches_reference_golden():
    g = load_golden('ddpm_cascade_dim32.pt')
    sdb, sds = synth_weights('test_base', g['wseed_base']), synth_weights('test_sr', g['wseed_sr'])
    cb = unet_ref.unet_config(**g['base_kwargs'])
    cs = unet_ref.unet_config(**g['sr_kwargs'], lowres_cond=True)
    torch.manual_seed(g['seed'])
    with torch.no_grad():
        outs = sampler_ref.imagen_sample([(sdb, cb), (sds, cs)], (16, 32), text_embeds=g['text_embeds'], timesteps=g['timesteps'],
                                         cond_scale=g['cond_scale'], return_all_unet_outputs=True)
    for o, r in zip(outs, g['outs']):
        assert (o - r).abs().max() < 1e-4


def test_edm_cascade_matches_reference_golden():
    g = load_golden('edm_cascade_dim32.pt')
    sdb, sds = synth_weights('test_base', g['wseed_base']), synth_weights('test_sr', g['wseed_sr'])
    cb = unet_ref.unet_config(**g['base_kwargs'])
    cs = unet_ref.unet_config(**g['sr_kwargs'], lowres_cond=True)
    torch.manual_seed(g['seed'])
    with torch.no_grad():
        out = sampler_ref.elucidated_sample([(sdb, cb), (sds, cs)], (16, 32), text_embeds=g['text_embeds'], cond_scale=g['cond_scale'],
                                            hparams=dict(num_sample_steps=g['num_sample_steps']))
    # sigma_max = 80 puts |x| ~ 80 in fp32; the untrained net amplifies 1-ulp re-association differences (see make_golden.py)
    assert (out - g['out']).abs().max() < 2e-3


def test_schedule_known_answers():
    g = load_golden('schedules.pt')
    assert torch.allclose(sampler_ref.alpha_cosine_log_snr(g['t']), g['cosine'], atol=1e-6)
    assert torch.allclose(sampler_ref.beta_linear_log_snr(g['t']), g['linear'], atol=1e-6)
    assert torch.allclose(sampler_ref.edm_sample_schedule(4, 7, 0.002, 80), g['edm_sigmas'], rtol=1e-6)
    # SURVEY.md appendix A.2 / A.3 probe values
    assert abs(sampler_ref.alpha_cosine_log_snr(torch.tensor([0.5])).item() + 0.0249) < 1e-3
    assert abs(sampler_ref.beta_linear_log_snr(torch.tensor([0.2])).item() - 0.7093) < 1e-3
    assert abs(sampler_ref.edm_sample_schedule(4, 7, 0.002, 80)[1].item() - 9.7232) < 1e-2
