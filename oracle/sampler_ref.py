"""ORACLE (test infrastructure, NOT product code) -- CPU fp32 restatement of the
reference's sampling loops: Imagen.p_sample_loop / Imagen.sample (DDPM, continuous
time) and ElucidatedImagen.one_unet_sample / .sample (EDM stochastic Heun).

Citations are to /root/reference/imagen_pytorch/{imagen_pytorch,elucidated_imagen}.py.
RNG draws are made in the reference's order and shapes so
that, on the same torch generator state, the trajectories coincide.
Parity status: PINNED against the live reference by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from math import sqrt
import torch
import torch.nn.functional as F
from torch.special import expm1

from . import unet_ref

# --------------------------------------------------------------------------- schedules


def beta_linear_log_snr(t):                           # imagen_pytorch.py:212-214
    return -torch.log(expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t, s: float = 0.008):        # imagen_pytorch.py:216-218
    return -torch.log(((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1).clamp(min=1e-5))


def log_snr_to_alpha_sigma(log_snr):                  # imagen_pytorch.py:220-221
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


LOG_SNR = {'linear': beta_linear_log_snr, 'cosine': alpha_cosine_log_snr}


def sampling_timesteps(num_timesteps, batch, device='cpu'):         # get_sampling_timesteps, imagen_pytorch.py:245-250
    times = torch.linspace(1., 0., num_timesteps + 1, device=device)
    times = times[None, :].expand(batch, -1)
    times = torch.stack((times[:, :-1], times[:, 1:]), dim=0)
    return times.unbind(dim=-1)


def dynamic_threshold(x_start, percentile=0.95):      # imagen_pytorch.py:2094-2105 / elucidated_imagen.py:309-321
    s = torch.quantile(x_start.flatten(1).abs(), percentile, dim=-1)
    s = s.clamp(min=1.)
    s = s.view(-1, *((1,) * (x_start.ndim - 1)))
    return x_start.clamp(-s, s) / s


def q_posterior(log_snr_fn, x_start, x_t, t, t_next):  # imagen_pytorch.py:252-270
    pad = lambda v: v.view(-1, *((1,) * (x_t.ndim - 1)))
    log_snr, log_snr_next = pad(log_snr_fn(t)), pad(log_snr_fn(t_next))
    alpha, sigma = log_snr_to_alpha_sigma(log_snr)
    alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
    c = -expm1(log_snr - log_snr_next)
    mean = alpha_next * (x_t * (1 - c) / alpha + c * x_start)
    var = (sigma_next ** 2) * c
    log_var = torch.log(var.clamp(min=1e-20))
    return mean, var, log_var


def resize_nearest(img, size):                        # resize_image_to, imagen_pytorch.py:152-168
    if img.shape[-1] == size:
        return img
    return F.interpolate(img, size, mode='nearest')


# --------------------------------------------------------------------------- DDPM


def q_sample(log_snr_fn, x_start, t, noise):           # imagen_pytorch.py:272-284
    log_snr = log_snr_fn(t).view(-1, *((1,) * (x_start.ndim - 1)))
    alpha, sigma = log_snr_to_alpha_sigma(log_snr)
    return alpha * x_start + sigma * noise


def q_sample_from_to(log_snr_fn, x_from, from_t, to_t, noise):   # imagen_pytorch.py:286-306
    pad = lambda v: v.view(-1, *((1,) * (x_from.ndim - 1)))
    alpha, sigma = log_snr_to_alpha_sigma(pad(log_snr_fn(from_t)))
    alpha_to, sigma_to = log_snr_to_alpha_sigma(pad(log_snr_fn(to_t)))
    return x_from * (alpha_to / alpha) + noise * (sigma_to * alpha - sigma * alpha_to) / alpha


def prepare_inpaint(inpaint_images, inpaint_masks, size, normalize=True):
    """imagen_pytorch.py:2216-2222 / elucidated_imagen.py:455-462: images to [-1, 1] and the stage resolution, masks (b, h, w) bool ->
    (b, 1, size, size) bool through a nearest resize of their float image."""
    imgs = inpaint_images * 2 - 1 if normalize else inpaint_images
    imgs = resize_nearest(imgs, size)
    masks = resize_nearest(inpaint_masks[:, None].float(), size).bool()
    return imgs, masks


def ddpm_p_sample_loop(unet_fn, shape, *, schedule='cosine', timesteps=1000, cond_scale=1.,
                       pred_objective='noise', dynamic_thresholding=True, percentile=0.95,
                       unet_kwargs=None, lowres_log_snr=None, trace=None, randn=torch.randn,
                       init_images=None, skip_steps=None, inpaint_images=None, inpaint_masks=None,
                       inpaint_resample_times=5, self_cond=False):
    """Imagen.p_sample_loop (imagen_pytorch.py:2167-2289) + p_sample (:2112-2165) +
    p_mean_variance (:2042-2110), incl. init image / skipped steps (:2205-2206, :2230-2231) and RePaint
    inpainting (:2216-2222, :2245-2279) and self-conditioning on the previous step's thresholded x_start (:2252).  inpaint_images / init_images arrive already
    normalised and resized (prepare_inpaint / the caller), as p_sample_loop's own callers hand them over.
    unet_fn(x, log_snr, cond_scale=..., **unet_kwargs) -> prediction."""
    unet_kwargs = dict(unet_kwargs or {})
    log_snr_fn = LOG_SNR[schedule]
    batch = shape[0]
    img = randn(shape)                                                         # :2195
    if init_images is not None:
        img = img + init_images                                                # :2205-2206
    has_inpainting = inpaint_images is not None and inpaint_masks is not None
    resample_times = inpaint_resample_times if has_inpainting else 1
    steps = list(sampling_timesteps(timesteps, batch, img.device))[(skip_steps or 0):]   # :2226-2231
    x_start = None                                                             # :2210 (self conditioning)
    for times, times_next in steps:                                            # :2242
        is_last_timestep = times_next == 0
        for r in reversed(range(resample_times)):
            if has_inpainting:                                                 # :2248-2250
                noised = q_sample(log_snr_fn, inpaint_images, times, randn(tuple(img.shape)))
                img = img * ~inpaint_masks + noised * inpaint_masks
            sc = dict(self_cond=x_start) if self_cond else {}                  # :2252
            pred = unet_fn(img, log_snr_fn(times), cond_scale=cond_scale,
                           lowres_noise_times=lowres_log_snr, **sc, **unet_kwargs)   # :2072-2083
            pad = lambda v: v.view(-1, 1, 1, 1)
            alpha, sigma = log_snr_to_alpha_sigma(pad(log_snr_fn(times)))
            if pred_objective == 'noise':                                      # :314-318
                x_start = (img - sigma * pred) / alpha.clamp(min=1e-8)
            elif pred_objective == 'x_start':
                x_start = pred
            elif pred_objective == 'v':                                        # :308-312
                x_start = alpha * img - sigma * pred
            else:
                raise ValueError(pred_objective)
            if dynamic_thresholding:
                x_start = dynamic_threshold(x_start, percentile)
            else:
                x_start = x_start.clamp(-1., 1.)
            mean, _, log_var = q_posterior(log_snr_fn, x_start, img, times, times_next)
            noise = randn(tuple(img.shape))                                    # :2160 (randn_like)
            nonzero = (1 - (times_next == 0).float()).view(batch, 1, 1, 1)
            img = mean + nonzero * (0.5 * log_var).exp() * noise               # :2164
            if has_inpainting and not (r == 0 or bool(torch.all(is_last_timestep))):   # :2271-2278
                renoised = q_sample_from_to(log_snr_fn, img, times_next, times, randn(tuple(img.shape)))
                img = torch.where(is_last_timestep.view(batch, 1, 1, 1), img, renoised)
            if trace is not None:
                trace.append(img.clone())
    img = img.clamp(-1., 1.)                                                   # :2281
    if has_inpainting:
        img = img * ~inpaint_masks + inpaint_images * inpaint_masks            # :2285-2286
    return (img + 1) * 0.5                                                     # :2288, :196-197


def imagen_sample(unets, image_sizes, *, text_embeds, text_masks=None, timesteps=1000, cond_scale=1.,
                  noise_schedules=('cosine',), lowres_sample_noise_level=0.2, dynamic_thresholding=True,
                  pred_objectives='noise', stop_at_unet_number=None, return_all_unet_outputs=False, trace=None,
                  randn=torch.randn, init_images=None, skip_steps=None, inpaint_images=None, inpaint_masks=None,
                  inpaint_resample_times=5, start_at_unet_number=1, start_image=None, cond_images=None):
    """Imagen.sample (imagen_pytorch.py:2291-2498), text_embeds path, no video.
    unets: list of (state_dict, cfg).  start_at_unet_number / start_image: the upscale-only entry (:2396-2403)."""
    n = len(unets)
    timesteps = unet_ref._tup(timesteps, n)
    cond_scale = unet_ref._tup(cond_scale, n)
    pred_objectives = unet_ref._tup(pred_objectives, n)
    dynamic_thresholding = unet_ref._tup(dynamic_thresholding, n)
    sched = tuple(noise_schedules)
    sched = sched + ('cosine',) * max(0, 2 - len(sched))                       # :1853-1855
    sched = sched + ('linear',) * max(0, n - len(sched))
    if text_masks is None:
        text_masks = torch.any(text_embeds != 0., dim=-1)                      # :2337
    batch = text_embeds.shape[0]
    init_images = [None if im is None else im * 2 - 1 for im in unet_ref._tup(init_images, n)]   # :2385-2386 (normalize_img)
    skip_steps = unet_ref._tup(skip_steps, n)
    outputs, img = [], None
    if start_at_unet_number > 1:                                               # :2396-2403
        assert start_image is not None, 'starting image or video must be supplied if only doing upscaling'
        img = resize_nearest(start_image, image_sizes[start_at_unet_number - 2])
    for i, ((sd, cfg), size) in enumerate(zip(unets, image_sizes)):
        if i + 1 < start_at_unet_number:                                       # :2412-2414
            continue
        kw = dict(text_embeds=text_embeds, text_mask=text_masks)
        if cond_images is not None:
            kw['cond_images'] = cond_images                                    # :2465 (the U-Net resizes it, :1559)
        lowres_log_snr = None
        opt = dict(skip_steps=skip_steps[i], inpaint_resample_times=inpaint_resample_times)
        if init_images[i] is not None:
            opt['init_images'] = resize_nearest(init_images[i], size)          # :2453-2454
        if inpaint_images is not None and inpaint_masks is not None:
            opt['inpaint_images'], opt['inpaint_masks'] = prepare_inpaint(inpaint_images, inpaint_masks, size)
        if cfg['lowres_cond']:                                                 # :2443-2449
            lt = torch.full((batch,), lowres_sample_noise_level, dtype=torch.float32)
            low = resize_nearest(img, size) * 2 - 1
            ls = beta_linear_log_snr(lt)
            a, s = log_snr_to_alpha_sigma(ls.view(-1, 1, 1, 1))
            low = a * low + s * randn(tuple(low.shape))                        # q_sample :272-284 (randn_like)
            kw['lowres_cond_img'] = low
            lowres_log_snr = beta_linear_log_snr(lt)                           # :2081
        fn = lambda x, t, cond_scale, lowres_noise_times=None, _sd=sd, _cfg=cfg, **k: \
            unet_ref.unet_forward_with_cond_scale(_sd, _cfg, x, t, cond_scale=cond_scale,
                                                  lowres_noise_times=lowres_noise_times, **k)
        img = ddpm_p_sample_loop(fn, (batch, cfg['channels'], size, size), schedule=sched[i],
                                 timesteps=timesteps[i], cond_scale=cond_scale[i],
                                 pred_objective=pred_objectives[i],
                                 dynamic_thresholding=dynamic_thresholding[i], unet_kwargs=kw,
                                 lowres_log_snr=lowres_log_snr, trace=trace, randn=randn, self_cond=cfg.get('self_cond', False), **opt)
        outputs.append(img)
        if stop_at_unet_number is not None and stop_at_unet_number == i + 1:
            break
    return outputs if return_all_unet_outputs else outputs[-1]


# --------------------------------------------------------------------------- EDM


def edm_sample_schedule(num_sample_steps, rho, sigma_min, sigma_max):   # elucidated_imagen.py:376-390
    N = num_sample_steps
    inv_rho = 1 / rho
    steps = torch.arange(N, dtype=torch.float32)
    sigmas = (sigma_max ** inv_rho + steps / (N - 1) * (sigma_min ** inv_rho - sigma_max ** inv_rho)) ** rho
    return F.pad(sigmas, (0, 1), value=0.)


def edm_precond_forward(unet_fn, x, sigma, *, sigma_data, dynamic_thresholding=True, percentile=0.95, **kw):
    """preconditioned_network_forward, elucidated_imagen.py:340-369 (clamp=True)."""
    batch = x.shape[0]
    sig = torch.full((batch,), sigma)
    ps = sig.view(-1, 1, 1, 1)
    c_in = 1 * (ps ** 2 + sigma_data ** 2) ** -0.5
    c_noise = torch.log(sig.clamp(min=1e-20)) * 0.25                           # log() helper elucidated_imagen.py:72-73 (eps 1e-20)
    c_skip = (sigma_data ** 2) / (ps ** 2 + sigma_data ** 2)
    c_out = ps * sigma_data * (sigma_data ** 2 + ps ** 2) ** -0.5
    net_out = unet_fn(c_in * x, c_noise, **kw)
    out = c_skip * x + c_out * net_out
    if dynamic_thresholding:
        return dynamic_threshold(out, percentile)
    return out.clamp(-1., 1.)


def edm_one_unet_sample(unet_fn, shape, *, num_sample_steps=32, sigma_min=0.002, sigma_max=80, sigma_data=0.5,
                        rho=7, S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003, cond_scale=1.,
                        dynamic_thresholding=True, unet_kwargs=None, trace=None, randn=torch.randn,
                        init_images=None, skip_steps=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5,
                        self_cond=False):
    """ElucidatedImagen.one_unet_sample, elucidated_imagen.py:392-545 (init image :446-447, skipped steps :476-477, RePaint
    inpainting :455-462, :498-499, :533-536, :541-542; inpaint / init images arrive normalised and resized)."""
    unet_kwargs = dict(unet_kwargs or {})
    sigmas = edm_sample_schedule(num_sample_steps, rho, sigma_min, sigma_max)
    gammas = torch.where((sigmas >= S_tmin) & (sigmas <= S_tmax),
                         min(S_churn / num_sample_steps, sqrt(2) - 1), 0.)
    images = sigmas[0] * randn(shape)                                          # :442
    if init_images is not None:
        images = images + init_images                                          # :446-447
    has_inpainting = inpaint_images is not None and inpaint_masks is not None
    resample_times = inpaint_resample_times if has_inpainting else 1
    kw = dict(sigma_data=sigma_data, dynamic_thresholding=dynamic_thresholding, cond_scale=cond_scale, **unet_kwargs)
    steps = list(zip(sigmas[:-1], sigmas[1:], gammas[:-1]))[(skip_steps or 0):]   # :476-477
    x_start = None                                                             # :451 (self conditioning)
    for ind, (sigma, sigma_next, gamma) in enumerate(steps):
        is_last_timestep = ind == len(steps) - 1
        sigma, sigma_next, gamma = (t.item() for t in (sigma, sigma_next, gamma))   # :484
        for r in reversed(range(resample_times)):
            eps = S_noise * randn(shape)                                       # :489
            sigma_hat = sigma + gamma * sigma
            added_noise = sqrt(sigma_hat ** 2 - sigma ** 2) * eps
            images_hat = images + added_noise
            if has_inpainting:                                                 # :498-499
                images_hat = images_hat * ~inpaint_masks + (inpaint_images + added_noise) * inpaint_masks
            sc = dict(self_cond=x_start) if self_cond else {}                  # :496
            model_output = edm_precond_forward(unet_fn, images_hat, sigma_hat, **sc, **kw)
            d = (images_hat - model_output) / sigma_hat
            images_next = images_hat + (sigma_next - sigma_hat) * d
            if sigma_next != 0:                                                # :515-529
                sc = dict(self_cond=model_output) if self_cond else {}         # :518
                model_output_next = edm_precond_forward(unet_fn, images_next, sigma_next, **sc, **kw)
                d_prime = (images_next - model_output_next) / sigma_next
                images_next = images_hat + 0.5 * (sigma_next - sigma_hat) * (d + d_prime)
            images = images_next
            if has_inpainting and not (r == 0 or is_last_timestep):           # :533-536
                images = images + (sigma - sigma_next) * randn(shape)
            x_start = model_output if sigma_next == 0 else model_output_next   # :538
            if trace is not None:
                trace.append(images.clone())
    images = images.clamp(-1., 1.)
    if has_inpainting:
        images = images * ~inpaint_masks + inpaint_images * inpaint_masks      # :541-542
    return (images + 1) * 0.5


def elucidated_sample(unets, image_sizes, *, text_embeds, text_masks=None, cond_scale=1.,
                      lowres_sample_noise_level=0.2, dynamic_thresholding=True, hparams=None, trace=None, randn=torch.randn,
                      init_images=None, skip_steps=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5,
                      return_all_unet_outputs=False):
    """ElucidatedImagen.sample (elucidated_imagen.py:547-751), text_embeds path."""
    n = len(unets)
    hparams = hparams or {}
    cond_scale = unet_ref._tup(cond_scale, n)
    if text_masks is None:
        text_masks = torch.any(text_embeds != 0., dim=-1)
    batch = text_embeds.shape[0]
    init_images = [None if im is None else im * 2 - 1 for im in unet_ref._tup(init_images, n)]   # :640-641
    skip_steps = unet_ref._tup(skip_steps, n)
    img, outputs = None, []
    for i, ((sd, cfg), size) in enumerate(zip(unets, image_sizes)):
        kw = dict(text_embeds=text_embeds, text_mask=text_masks)
        opt = dict(skip_steps=skip_steps[i], inpaint_resample_times=inpaint_resample_times)
        if init_images[i] is not None:
            opt['init_images'] = resize_nearest(init_images[i], size)          # :709-710
        if inpaint_images is not None and inpaint_masks is not None:
            opt['inpaint_images'], opt['inpaint_masks'] = prepare_inpaint(inpaint_images, inpaint_masks, size)
        if cfg['lowres_cond']:                                                 # :699-705
            lt = torch.full((batch,), lowres_sample_noise_level, dtype=torch.float32)
            low = resize_nearest(img, size) * 2 - 1
            a, s = log_snr_to_alpha_sigma(beta_linear_log_snr(lt).view(-1, 1, 1, 1))
            kw['lowres_cond_img'] = a * low + s * randn(tuple(low.shape))
            kw['lowres_noise_times'] = lt                                      # raw times, NOT log-snr (:700, passed through **kwargs :727-728)
        hp = {k: (unet_ref._tup(v, n)[i]) for k, v in hparams.items()}
        fn = lambda x, t, cond_scale, _sd=sd, _cfg=cfg, **k: \
            unet_ref.unet_forward_with_cond_scale(_sd, _cfg, x, t, cond_scale=cond_scale, **k)
        img = edm_one_unet_sample(fn, (batch, cfg['channels'], size, size), cond_scale=cond_scale[i],
                                  dynamic_thresholding=dynamic_thresholding, unet_kwargs=kw, trace=trace, randn=randn,
                                  self_cond=cfg.get('self_cond', False), **hp, **opt)
        outputs.append(img)
    return outputs if return_all_unet_outputs else img
