"""``Imagen``: cascaded continuous-time DDPM sampler with the reference's constructor and
``.sample()`` API (imagen_pytorch.py:1787-2498), driving ``UnetPlan`` launch plans.

Per cascade stage the host does, once: build the per-step coefficient table with the same torch
ops as GaussianDiffusionContinuousTimes (so the scalars are bit-identical to the reference's),
hoist the conditioning (``UnetPlan.prepare``), draw the initial noise with torch's generator in the
reference's order -- and then replays ONE captured CUDA graph per denoising step:
    randn_like -> [~350 kernel launches of the batched cond+null U-Net] -> fused DDPM step kernel.
Training (``forward`` / ``p_losses``) is out of scope of this hot-path implementation.
"""
from __future__ import annotations

import math
import os
from contextlib import nullcontext
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn
from torch.special import expm1

from . import _lib
from .params import cast_tuple
from .unet import Unet, NullUnet

from .t5 import DEFAULT_T5_NAME, get_encoded_dim, t5_encode_text, TextEmbedCache   # noqa: E402  (reference: imagen_pytorch.py:23)


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else (d() if callable(d) else d)


def pad_tuple_to_length(t, length, fillvalue=None):
    return t if len(t) >= length else (*t, *((fillvalue,) * (length - len(t))))


def resize_image_to(image, target_image_size, clamp_range=None, mode='nearest'):   # :152-168
    if image.shape[-1] == target_image_size:
        return image
    out = F.interpolate(image, target_image_size, mode=mode)
    if exists(clamp_range):
        out = out.clamp(*clamp_range)
    return out


def normalize_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_zero_to_one(img):
    return (img + 1) * 0.5


def _log(t, eps=1e-12):
    return torch.log(t.clamp(min=eps))


def beta_linear_log_snr(t):                                                    # :212-214
    return -torch.log(expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t, s: float = 0.008):                                 # :216-218
    return -_log((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def log_snr_to_alpha_sigma(log_snr):                                           # :220-221
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


class GaussianDiffusionContinuousTimes(nn.Module):
    """Host-side schedule (imagen_pytorch.py:223-318); only the sampling members."""

    def __init__(self, *, noise_schedule, timesteps=1000):
        super().__init__()
        if noise_schedule == 'linear':
            self.log_snr = beta_linear_log_snr
        elif noise_schedule == 'cosine':
            self.log_snr = alpha_cosine_log_snr
        else:
            raise ValueError(f'invalid noise schedule {noise_schedule}')
        self.num_timesteps = timesteps

    def get_times(self, batch_size, noise_level, *, device):
        return torch.full((batch_size,), noise_level, device=device, dtype=torch.float32)

    def get_condition(self, times):
        return self.log_snr(times) if exists(times) else None

    def get_sampling_timesteps(self, batch, *, device):
        times = torch.linspace(1., 0., self.num_timesteps + 1, device=device)
        times = times[None, :].expand(batch, -1)
        times = torch.stack((times[:, :-1], times[:, 1:]), dim=0)
        return times.unbind(dim=-1)

    def q_sample(self, x_start, t, noise=None):                               # :272-284
        if isinstance(t, float):
            t = torch.full((x_start.shape[0],), t, device=x_start.device, dtype=x_start.dtype)
        noise = default(noise, lambda: torch.randn_like(x_start))
        log_snr = self.log_snr(t).type(x_start.dtype)
        alpha, sigma = log_snr_to_alpha_sigma(log_snr.view(-1, *((1,) * (x_start.ndim - 1))))
        return alpha * x_start + sigma * noise, log_snr, alpha, sigma

    def ddpm_coefficients(self, device):
        """Per-step scalars of p_mean_variance / q_posterior (:252-270, :314-318), [T, 8] fp32, evaluated with
        the reference's torch ops on `device` so they round identically."""
        times = torch.linspace(1., 0., self.num_timesteps + 1, device=device)
        t, t_next = times[:-1], times[1:]
        log_snr, log_snr_next = self.log_snr(t), self.log_snr(t_next)
        alpha, sigma = log_snr_to_alpha_sigma(log_snr)
        alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
        c = -expm1(log_snr - log_snr_next)
        log_var = _log((sigma_next ** 2) * c, eps=1e-20)
        nonzero = 1 - (t_next == 0).float()
        noise_std = nonzero * (0.5 * log_var).exp()
        z = torch.zeros_like(alpha)
        coefs = torch.stack((sigma, alpha, 1 / alpha.clamp(min=1e-8), alpha_next, c, noise_std, z, z), dim=1).contiguous()
        return coefs, log_snr


    def repaint_coefficients(self, device):
        """Per-step scalars of the RePaint inpainting glue, [T, 5] fp32 = (alpha_t, sigma_t) of q_sample (:272-284) and
        (c1, c2, alpha_from) of q_sample_from_to(x, t_next -> t) (:286-306), with the reference's torch ops on `device`."""
        times = torch.linspace(1., 0., self.num_timesteps + 1, device=device)
        t, t_next = times[:-1], times[1:]
        alpha, sigma = log_snr_to_alpha_sigma(self.log_snr(t))                 # "to" of the re-noising, and q_sample's time
        alpha_f, sigma_f = log_snr_to_alpha_sigma(self.log_snr(t_next))        # "from"
        return torch.stack((alpha, sigma, alpha / alpha_f, sigma * alpha_f - sigma_f * alpha, alpha_f), dim=1).contiguous()


def quantile_ranks(n, q, device):
    """(lower rank, upper rank, lerp weight) exactly as torch.quantile's linear interpolation computes them."""
    rank = torch.tensor(q, dtype=torch.float32, device=device) * (n - 1)
    lo, hi = rank.floor(), rank.ceil()
    return int(lo.item()), int(hi.item()), float((rank - lo).item())


class _SamplerBase(nn.Module):
    """Parts shared by Imagen and ElucidatedImagen: cascade bookkeeping (imagen_pytorch.py:1886-1975)."""

    def _init_common(self, unets, *, image_sizes, text_encoder_name, text_embed_dim, channels, cond_drop_prob, condition_on_text,
                     auto_normalize_img, dynamic_thresholding, dynamic_thresholding_percentile, lowres_noise_schedule,
                     lowres_sample_noise_level, resize_mode, random_crop_sizes=None):
        self.condition_on_text = condition_on_text
        self.unconditional = not condition_on_text
        self.channels = channels
        unets = cast_tuple(unets)
        num_unets = len(unets)
        self.random_crop_sizes = cast_tuple(random_crop_sizes, num_unets)
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(noise_schedule=lowres_noise_schedule)
        self.text_encoder_name = text_encoder_name
        self.text_embed_dim = default(text_embed_dim, lambda: get_encoded_dim(text_encoder_name))
        self.encode_text = partial(t5_encode_text, name=text_encoder_name)    # :1884; replaceable (e.g. enable_text_embed_cache)
        self.unets = nn.ModuleList([])
        for ind, one_unet in enumerate(unets):
            assert isinstance(one_unet, (Unet, NullUnet))
            one_unet = one_unet.cast_model_parameters(
                lowres_cond=not ind == 0, cond_on_text=self.condition_on_text,
                text_embed_dim=self.text_embed_dim if self.condition_on_text else None,
                channels=self.channels, channels_out=self.channels)
            self.unets.append(one_unet)
        image_sizes = cast_tuple(image_sizes)
        self.image_sizes = image_sizes
        assert num_unets == len(image_sizes), f'you did not supply the correct number of u-nets ({len(unets)}) for resolutions {image_sizes}'
        self.sample_channels = cast_tuple(self.channels, num_unets)
        self.is_video = False
        self.resize_to = lambda img, size: resize_image_to(img, size, mode=resize_mode)
        lowres_conditions = tuple(map(lambda t: t.lowres_cond, self.unets))
        assert lowres_conditions == (False, *((True,) * (num_unets - 1))), \
            'the first unet must be unconditioned (by low resolution image), and the rest of the unets must have `lowres_cond` set to True'
        self.lowres_sample_noise_level = lowres_sample_noise_level
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.
        self.normalize_img = normalize_neg_one_to_one if auto_normalize_img else (lambda t: t)
        self.unnormalize_img = unnormalize_zero_to_one if auto_normalize_img else (lambda t: t)
        self.auto_normalize_img = auto_normalize_img
        self.input_image_range = (0. if auto_normalize_img else -1., 1.)
        self.dynamic_thresholding = cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile
        self.register_buffer('_temp', torch.tensor([0.]), persistent=False)
        self.to(next(self.unets.parameters()).device)
        self.last_launch_count = 0

    @property
    def device(self):
        return self._temp.device

    def enable_text_embed_cache(self, capacity=4096):
        """Route sample(texts=...) through an LRU of per-prompt embeddings (t5.TextEmbedCache): repeated prompts skip the encoder."""
        self.text_embed_cache = TextEmbedCache(capacity, encode_fn=self.encode_text)
        self.encode_text = partial(self.text_embed_cache.encode, name=self.text_encoder_name)
        return self.text_embed_cache

    def _encode_texts(self, texts, text_embeds, text_masks, device):
        """texts -> (text_embeds, text_masks) when no embeddings were passed (imagen_pytorch.py:2326-2332)."""
        if exists(texts) and not exists(text_embeds) and not self.unconditional:
            assert all([*map(len, texts)]), 'text cannot be empty'
            text_embeds, text_masks = self.encode_text(texts, return_attn_mask=True)
            text_embeds, text_masks = text_embeds.to(device), text_masks.to(device)
        return text_embeds, text_masks

    def force_unconditional_(self):
        self.condition_on_text = False
        self.unconditional = True
        for unet in self.unets:
            unet.cond_on_text = False

    def get_unet(self, unet_number):                                           # B200 has 180 GB: every U-Net stays resident (:1988-2002)
        assert 0 < unet_number <= len(self.unets)
        return self.unets[unet_number - 1]

    def reset_unets_all_one_device(self, device=None):
        self.unets.to(default(device, self.device))

    def one_unet_in_gpu(self, unet_number=None, unet=None):                    # accepted and ignored (:2011-2028)
        return nullcontext()

    def _check_sample_args(self, texts, text_embeds, text_masks, unsupported):
        for name, val in unsupported.items():
            if exists(val):
                raise NotImplementedError(f'sample({name}=...) is outside the B200 sampling hot path (see DESIGN.md)')
        if not self.unconditional:
            assert exists(text_embeds), 'text must be passed in if the network was not trained without text `condition_on_text` must be set to `False` when training'
        assert not (self.condition_on_text and not exists(text_embeds)), 'text or text encodings must be passed into imagen if specified'
        assert not (not self.condition_on_text and exists(text_embeds)), 'imagen specified not to be conditioned on text, yet it is presented'
        assert not (exists(text_embeds) and text_embeds.shape[-1] != self.text_embed_dim), \
            f'invalid text embedding dimension being passed in (should be {self.text_embed_dim})'

    def _check_cond_images(self, cond_images, start_at_unet_number=1, stop_at_unet_number=None):
        """Unet.forward's check (imagen_pytorch.py:1555), raised before any device work."""
        for i, unet in enumerate(self.unets, 1):
            if i < start_at_unet_number or (exists(stop_at_unet_number) and i > stop_at_unet_number) or isinstance(unet, NullUnet):
                continue
            assert not (unet.has_cond_image ^ exists(cond_images)), \
                'you either requested to condition on an image on the unet, but the conditioning image is not supplied, or vice versa'

    def _lowres_conditioning(self, img, image_size, batch_size, level, device):
        """Noised low-res conditioning of a cascade stage (:2443-2449 / elucidated_imagen.py:699-705)."""
        times = self.lowres_noise_schedule.get_times(batch_size, level, device=device)
        low = self.resize_to(img, image_size)
        low = self.normalize_img(low)
        low, *_ = self.lowres_noise_schedule.q_sample(x_start=low, t=times, noise=torch.randn_like(low))
        return low, times

    def forward(self, *args, **kwargs):
        raise NotImplementedError('training (p_losses / forward) is outside the B200 sampling hot path; train with the reference and '
                                  'load the state_dict here')


def require_cuda(device):
    if torch.device(device).type != 'cuda':
        raise _lib.B200Error(f'imagen_pytorch_b200 samples on an sm_100 CUDA device only (module is on {device}); there is no CPU / torch '
                             f'fallback -- move the module to cuda, or use the reference implementation on CPU')
    _lib.require_device(torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())


def _use_graph():
    return os.environ.get('B200_IMAGEN_NO_GRAPH', '0') != '1'


class Imagen(_SamplerBase):
    def __init__(self, unets, *, image_sizes, text_encoder_name=DEFAULT_T5_NAME, text_embed_dim=None, channels=3, timesteps=1000,
                 cond_drop_prob=0.1, loss_type='l2', noise_schedules='cosine', pred_objectives='noise', random_crop_sizes=None,
                 lowres_noise_schedule='linear', lowres_sample_noise_level=0.2, per_sample_random_aug_noise_level=False,
                 condition_on_text=True, auto_normalize_img=True, dynamic_thresholding=True, dynamic_thresholding_percentile=0.95,
                 only_train_unet_number=None, temporal_downsample_factor=1, resize_cond_video_frames=True, resize_mode='nearest',
                 min_snr_loss_weight=True, min_snr_gamma=5):
        super().__init__()
        if loss_type not in ('l1', 'l2', 'huber'):
            raise NotImplementedError()
        self.loss_type = loss_type
        num_unets = len(cast_tuple(unets))
        timesteps = cast_tuple(timesteps, num_unets)
        noise_schedules = cast_tuple(noise_schedules)
        noise_schedules = pad_tuple_to_length(noise_schedules, 2, 'cosine')    # :1853-1855
        noise_schedules = pad_tuple_to_length(noise_schedules, num_unets, 'linear')
        self.noise_schedulers = nn.ModuleList([GaussianDiffusionContinuousTimes(noise_schedule=s, timesteps=t)
                                               for t, s in zip(timesteps, noise_schedules)])
        self.pred_objectives = cast_tuple(pred_objectives, num_unets)
        self.per_sample_random_aug_noise_level = per_sample_random_aug_noise_level
        self.only_train_unet_number = only_train_unet_number
        self._init_common(unets, image_sizes=image_sizes, text_encoder_name=text_encoder_name, text_embed_dim=text_embed_dim,
                          channels=channels, cond_drop_prob=cond_drop_prob, condition_on_text=condition_on_text,
                          auto_normalize_img=auto_normalize_img, dynamic_thresholding=dynamic_thresholding,
                          dynamic_thresholding_percentile=dynamic_thresholding_percentile, lowres_noise_schedule=lowres_noise_schedule,
                          lowres_sample_noise_level=lowres_sample_noise_level, resize_mode=resize_mode, random_crop_sizes=random_crop_sizes)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def p_sample_loop(self, unet, shape, *, noise_scheduler, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None,
                      text_mask=None, cond_images=None, inpaint_images=None, inpaint_masks=None, init_images=None, skip_steps=None,
                      cond_scale=1, pred_objective='noise', dynamic_threshold=True, use_tqdm=True, **unsupported):
        """Imagen.p_sample_loop (:2167-2289): T ancestral steps; one CUDA-graph replay per step."""
        inpaint_resample_times = unsupported.pop('inpaint_resample_times', 5)
        inpaint_images = default(unsupported.pop('inpaint_videos', None), inpaint_images)
        for name, val in unsupported.items():
            if exists(val):
                raise NotImplementedError(f'p_sample_loop({name}=...) is outside the B200 sampling hot path')
        assert not (cond_scale != 1. and not self.can_classifier_guidance), \
            'imagen was not trained with conditional dropout, and thus one cannot use classifier free guidance (cond_scale anything other than 1)'
        device = self.device
        require_cuda(device)
        B, Cimg, H, W = shape
        T = noise_scheduler.num_timesteps
        R = B if cond_scale == 1 else 2 * B
        objective = {'noise': 0, 'x_start': 1, 'v': 2}.get(pred_objective)
        if objective is None:
            raise ValueError(f'unknown objective {pred_objective}')
        with torch.cuda.device(device):
            plan = unet.plan(R, B, H, W, T, device)
            coefs, log_snr = noise_scheduler.ddpm_coefficients(device)
            keep = torch.cat((torch.ones(B, dtype=torch.bool, device=device), torch.zeros(R - B, dtype=torch.bool, device=device)))
            plan.prepare(log_snr, text_embeds=text_embeds, text_mask=text_mask, keep=keep, lowres_cond_img=lowres_cond_img,
                         lowres_noise_times=self.lowres_noise_schedule.get_condition(lowres_noise_times), cond_images=cond_images)
            chw = Cimg * H * W
            q_lo, q_hi, q_w = quantile_ranks(chw, self.dynamic_thresholding_percentile, device)
            x = plan.x_in
            x.copy_(torch.randn(shape, device=device))                         # :2195
            if exists(init_images):
                x.add_(init_images.to(device=device, dtype=torch.float32))     # :2205-2206
            skip = default(skip_steps, 0)                                      # :2230-2231
            assert 0 <= skip <= T
            if skip:
                plan.slots.fill_(skip)                                         # the device-side schedule slot the first step reads
            has_inpainting = exists(inpaint_images) and exists(inpaint_masks)
            if has_inpainting:                                                 # :2216-2222
                known = self.resize_to(self.normalize_img(inpaint_images.to(device=device, dtype=torch.float32)), W).contiguous()
                mask = self.resize_to(inpaint_masks.to(device)[:, None].float(), W).bool().to(torch.uint8).contiguous()
                assert known.shape == tuple(shape) and mask.shape == (B, 1, H, W)
            # persistent per-plan sampler buffers: the captured step graph bakes their addresses and is reused by later sample() calls
            st_ = plan.sampler_state.setdefault('ddpm', {})
            if 'noise' not in st_:
                st_['noise'] = torch.empty(shape, dtype=torch.float32, device=device)
                st_['coefs'] = torch.empty((plan.S, 8), dtype=torch.float32, device=device)
                st_['graphs'] = {}
            noise, coef_buf = st_['noise'], st_['coefs']
            coef_buf[:T].copy_(coefs)
            lib = plan.lib
            key = (float(cond_scale), objective, bool(dynamic_threshold), q_lo, q_hi, q_w)
            sc_ptr = None
            if unet.self_cond:                                                 # x_start of the previous step conditions the next (:2210, :2252)
                plan.sc_in.zero_()
                sc_ptr = plan.sc_in.data_ptr()

            def one_step():
                noise.copy_(torch.randn_like(x))                               # :2160 (drawn every step, also the last)
                plan.launch()
                st = torch.cuda.current_stream(device).cuda_stream
                _lib.check(lib.b200_ddpm_step_sc(x.data_ptr(), plan.pred.data_ptr(), noise.data_ptr(), coef_buf.data_ptr(), plan.slots.data_ptr(),
                                                 R, B, chw, float(cond_scale), objective, int(bool(dynamic_threshold)), q_lo, q_hi, q_w,
                                                 sc_ptr, st), 'b200_ddpm_step')

            out = torch.empty(shape, dtype=torch.float32, device=device)
            if not has_inpainting:
                self.last_launch_count = self._run_steps(one_step, T - skip, plan, device, use_tqdm, launches_per_step=plan.n_launches + 2,
                                                         graph_cache=st_['graphs'], graph_key=key)
                _lib.check(lib.b200_finalize_images(x.data_ptr(), out.data_ptr(), out.numel(), int(self.auto_normalize_img),
                                                    torch.cuda.current_stream(device).cuda_stream), 'b200_finalize_images')   # :2281, :2288
            else:
                # RePaint (:2245-2279): every timestep is resampled `inpaint_resample_times` times at the SAME schedule slot, so the
                # loop is driven from the host (eager launches; the slot is rewritten before every network evaluation)
                rp = noise_scheduler.repaint_coefficients(device).tolist()
                st = torch.cuda.current_stream(device).cuda_stream
                launches = 0
                for i in range(skip, T):
                    a_t, s_t, c1, c2, a_from = rp[i]
                    last_t = i == T - 1                                        # times_next == 0
                    for r in reversed(range(inpaint_resample_times)):
                        qn = torch.randn_like(x)                               # q_sample's randn_like (:280)
                        _lib.check(lib.b200_inpaint_mix(x.data_ptr(), known.data_ptr(), mask.data_ptr(), qn.data_ptr(), a_t, s_t, B, Cimg,
                                                        H * W, st), 'b200_inpaint_mix')
                        plan.slots.fill_(i)
                        one_step()
                        launches += plan.n_launches + 4
                        if not (r == 0 or last_t):                             # :2271-2278
                            rn = torch.randn_like(x)
                            _lib.check(lib.b200_renoise(x.data_ptr(), rn.data_ptr(), c1, c2, a_from, x.numel(), st), 'b200_renoise')
                            launches += 2
                self.last_launch_count = launches
                _lib.check(lib.b200_finalize_images(x.data_ptr(), x.data_ptr(), x.numel(), 0, st), 'b200_finalize_images')     # :2281 clamp
                _lib.check(lib.b200_inpaint_mix(x.data_ptr(), known.data_ptr(), mask.data_ptr(), None, 1.0, 0.0, B, Cimg, H * W, st),
                           'b200_inpaint_mix')                                 # :2285-2286
                _lib.check(lib.b200_finalize_images(x.data_ptr(), out.data_ptr(), out.numel(), 2 | int(self.auto_normalize_img), st),
                           'b200_finalize_images')                             # :2288
        return out

    @staticmethod
    def _run_steps(one_step, n_steps, plan, device, use_tqdm, launches_per_step, graph_cache=None, graph_key=None):
        it = range(n_steps)
        if use_tqdm:
            try:
                from tqdm.auto import tqdm
                it = tqdm(it, desc='sampling loop time step', total=n_steps)
            except ImportError:
                pass
        if _use_graph() and n_steps > 1:
            graph = graph_cache.get(graph_key) if graph_cache is not None else None
            if graph is None:
                plan.launch()                                                  # warm-up outside capture (lazy function attributes); touches no RNG
                torch.cuda.synchronize(device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    one_step()
                if graph_cache is not None:
                    graph_cache[graph_key] = graph
            for _ in it:
                graph.replay()
        else:
            for _ in it:
                one_step()
        return launches_per_step * n_steps

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, texts=None, text_masks=None, text_embeds=None, video_frames=None, cond_images=None, cond_video_frames=None,
               post_cond_video_frames=None, inpaint_videos=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5,
               init_images=None, skip_steps=None, batch_size=1, cond_scale=1., lowres_sample_noise_level=None, start_at_unet_number=1,
               start_image_or_video=None, stop_at_unet_number=None, return_all_unet_outputs=False, return_pil_images=False, device=None,
               use_tqdm=True, use_one_unet_in_gpu=True):
        """Imagen.sample (:2291-2498).  `use_one_unet_in_gpu` is accepted and ignored (all U-Nets stay resident)."""
        was_training = self.training
        self.eval()
        try:
            return self._sample(texts, text_masks, text_embeds, dict(video_frames=video_frames,
                                cond_video_frames=cond_video_frames, post_cond_video_frames=post_cond_video_frames, inpaint_videos=inpaint_videos),
                                batch_size, cond_scale, lowres_sample_noise_level, start_at_unet_number, start_image_or_video,
                                stop_at_unet_number, return_all_unet_outputs, return_pil_images, device, use_tqdm,
                                options=dict(inpaint_images=inpaint_images, inpaint_masks=inpaint_masks, inpaint_resample_times=inpaint_resample_times,
                                             init_images=init_images, skip_steps=skip_steps, cond_images=cond_images))
        finally:
            self.train(was_training)

    def _sample(self, texts, text_masks, text_embeds, unsupported, batch_size, cond_scale, lowres_sample_noise_level, start_at_unet_number,
                start_image_or_video, stop_at_unet_number, return_all_unet_outputs, return_pil_images, device, use_tqdm, options=None):
        options = dict(options or {})
        device = default(device, self.device)
        self.reset_unets_all_one_device(device=device)
        text_embeds, text_masks = self._encode_texts(texts, text_embeds, text_masks, device)
        cond_images = options.get('cond_images')
        if exists(cond_images) and cond_images.dtype == torch.uint8:          # cast_uint8_images_to_float (:91-94, :2324)
            cond_images = cond_images / 255
        self._check_cond_images(cond_images, start_at_unet_number, stop_at_unet_number)
        self._check_sample_args(texts, text_embeds, text_masks, unsupported)
        if return_pil_images:
            raise NotImplementedError('return_pil_images: convert the returned tensor yourself')
        device = next(self.parameters()).device
        if not self.unconditional:
            text_embeds = text_embeds.to(device)
            text_masks = default(text_masks, lambda: torch.any(text_embeds != 0., dim=-1))   # :2337
            text_masks = text_masks.to(device)
            batch_size = text_embeds.shape[0]
        lowres_sample_noise_level = default(lowres_sample_noise_level, self.lowres_sample_noise_level)
        num_unets = len(self.unets)
        cond_scale = cast_tuple(cond_scale, num_unets)
        inpaint_images, inpaint_masks = options.get('inpaint_images'), options.get('inpaint_masks')
        if exists(inpaint_images):                                             # :2343-2350
            if self.unconditional:
                batch_size = inpaint_images.shape[0]
            assert inpaint_images.shape[0] == batch_size, 'number of inpainting images must be equal to the specified batch size on sample'
        init_images = cast_tuple(options.get('init_images'), num_unets)        # :2385-2388
        init_images = [self.normalize_img(im) if exists(im) else None for im in init_images]
        skip_steps = cast_tuple(options.get('skip_steps'), num_unets)
        img = None
        if start_at_unet_number > 1:                                           # :2396-2403
            assert start_at_unet_number <= num_unets, 'must start a unet that is less than the total number of unets'
            assert not exists(stop_at_unet_number) or start_at_unet_number <= stop_at_unet_number
            assert exists(start_image_or_video), 'starting image or video must be supplied if only doing upscaling'
            img = self.resize_to(start_image_or_video.to(device), self.image_sizes[start_at_unet_number - 2])
        outputs, launches = [], 0
        for unet_number, unet, channel, image_size, noise_scheduler, pred_objective, dynamic_threshold, unet_cond_scale, unet_init_images, \
                unet_skip_steps in zip(range(1, num_unets + 1), self.unets, self.sample_channels, self.image_sizes, self.noise_schedulers,
                                       self.pred_objectives, self.dynamic_thresholding, cond_scale, init_images, skip_steps):
            if unet_number < start_at_unet_number:
                continue
            assert not isinstance(unet, NullUnet), 'one cannot sample from null / placeholder unets'
            lowres_cond_img = lowres_noise_times = None
            if unet.lowres_cond:
                lowres_cond_img, lowres_noise_times = self._lowres_conditioning(img, image_size, batch_size, lowres_sample_noise_level, device)
            shape = (batch_size, self.channels, image_size, image_size)
            if exists(unet_init_images):
                unet_init_images = self.resize_to(unet_init_images.to(device), image_size)   # :2453-2454
            img = self.p_sample_loop(unet, shape, text_embeds=text_embeds, text_mask=text_masks, cond_scale=unet_cond_scale,
                                     lowres_cond_img=lowres_cond_img, lowres_noise_times=lowres_noise_times, noise_scheduler=noise_scheduler,
                                     pred_objective=pred_objective, dynamic_threshold=dynamic_threshold, use_tqdm=use_tqdm,
                                     inpaint_images=inpaint_images, inpaint_masks=inpaint_masks,
                                     inpaint_resample_times=options.get('inpaint_resample_times', 5), init_images=unet_init_images,
                                     skip_steps=unet_skip_steps, cond_images=cond_images)
            launches += self.last_launch_count
            outputs.append(img)
            if exists(stop_at_unet_number) and stop_at_unet_number == unet_number:
                break
        self.last_launch_count = launches
        return outputs if return_all_unet_outputs else outputs[-1]
