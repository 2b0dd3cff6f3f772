"""``ElucidatedImagen``: Karras et al. EDM stochastic Heun sampler around the same ``Unet``
(elucidated_imagen.py:77-751), with the reference's constructor and ``.sample()`` API.

All per-step scalars are data independent, so they are tabulated on the host exactly the way the
reference computes them (python doubles after ``.item()`` for sigma / gamma arithmetic, fp32 torch ops
for the preconditioning) and the three host syncs per step of the reference (:484) disappear: a step
is one CUDA-graph replay  [randn -> phase0 -> U-Net -> phase1 -> U-Net -> phase2].
"""
from __future__ import annotations

from collections import namedtuple
from math import sqrt

import torch
import torch.nn.functional as F

from . import _lib
from .imagen import (_SamplerBase, _use_graph, require_cuda, DEFAULT_T5_NAME, cast_tuple, default, exists, quantile_ranks)
from .unet import NullUnet

Hparams_fields = ['num_sample_steps', 'sigma_min', 'sigma_max', 'sigma_data', 'rho', 'P_mean', 'P_std', 'S_churn', 'S_tmin', 'S_tmax', 'S_noise']
Hparams = namedtuple('Hparams', Hparams_fields)


class ElucidatedImagen(_SamplerBase):
    def __init__(self, unets, *, image_sizes, text_encoder_name=DEFAULT_T5_NAME, text_embed_dim=None, channels=3, cond_drop_prob=0.1,
                 random_crop_sizes=None, resize_mode='nearest', temporal_downsample_factor=1, resize_cond_video_frames=True,
                 lowres_sample_noise_level=0.2, per_sample_random_aug_noise_level=False, condition_on_text=True, auto_normalize_img=True,
                 dynamic_thresholding=True, dynamic_thresholding_percentile=0.95, only_train_unet_number=None, lowres_noise_schedule='linear',
                 num_sample_steps=32, sigma_min=0.002, sigma_max=80, sigma_data=0.5, rho=7, P_mean=-1.2, P_std=1.2, S_churn=80,
                 S_tmin=0.05, S_tmax=50, S_noise=1.003):
        super().__init__()
        self.only_train_unet_number = only_train_unet_number
        self.per_sample_random_aug_noise_level = per_sample_random_aug_noise_level
        self._init_common(unets, image_sizes=image_sizes, text_encoder_name=text_encoder_name, text_embed_dim=text_embed_dim,
                          channels=channels, cond_drop_prob=cond_drop_prob, condition_on_text=condition_on_text,
                          auto_normalize_img=auto_normalize_img, dynamic_thresholding=dynamic_thresholding,
                          dynamic_thresholding_percentile=dynamic_thresholding_percentile, lowres_noise_schedule=lowres_noise_schedule,
                          lowres_sample_noise_level=lowres_sample_noise_level, resize_mode=resize_mode, random_crop_sizes=random_crop_sizes)
        num_unets = len(self.unets)
        hparams = [num_sample_steps, sigma_min, sigma_max, sigma_data, rho, P_mean, P_std, S_churn, S_tmin, S_tmax, S_noise]
        hparams = [cast_tuple(hp, num_unets) for hp in hparams]
        self.hparams = [Hparams(*unet_hp) for unet_hp in zip(*hparams)]

    # ---- schedule (elucidated_imagen.py:376-390) and per-step scalar table -------------------------------
    def sample_schedule(self, num_sample_steps, rho, sigma_min, sigma_max):
        N = num_sample_steps
        inv_rho = 1 / rho
        steps = torch.arange(num_sample_steps, device=self.device, dtype=torch.float32)
        sigmas = (sigma_max ** inv_rho + steps / (N - 1) * (sigma_min ** inv_rho - sigma_max ** inv_rho)) ** rho
        return F.pad(sigmas, (0, 1), value=0.)

    def _edm_tables(self, hp, sigma_min, sigma_max, device):
        sigmas = self.sample_schedule(hp.num_sample_steps, hp.rho, sigma_min, sigma_max)
        gammas = torch.where((sigmas >= hp.S_tmin) & (sigmas <= hp.S_tmax), min(hp.S_churn / hp.num_sample_steps, sqrt(2) - 1), 0.)
        sig, gam = sigmas.tolist(), gammas.tolist()                            # python doubles of the fp32 values == .item() (:484)
        sd = hp.sigma_data

        def precond(s):                                                        # c_in / c_skip / c_out on fp32 tensors (:325-335, :353-364)
            t = torch.full((1,), s, device=device)
            c_in = 1 * (t ** 2 + sd ** 2) ** -0.5
            c_skip = (sd ** 2) / (t ** 2 + sd ** 2)
            c_out = t * sd * (sd ** 2 + t ** 2) ** -0.5
            c_noise = torch.log(t.clamp(min=1e-20)) * 0.25                      # (:72-73, :334-335)
            return c_in.item(), c_skip.item(), c_out.item(), c_noise

        rows, times = [], []
        for i in range(hp.num_sample_steps):
            sigma, sigma_next, gamma = sig[i], sig[i + 1], gam[i]
            sigma_hat = sigma + gamma * sigma
            ci, cs, co, cn = precond(sigma_hat)
            times.append(cn)
            second = sigma_next != 0
            if second:
                ci2, cs2, co2, cn2 = precond(sigma_next)
                times.append(cn2)
            else:
                ci2 = cs2 = co2 = 0.
            rows.append([hp.S_noise, sqrt(sigma_hat ** 2 - sigma ** 2), sigma_hat, sigma_next, sigma_next - sigma_hat,
                         0.5 * (sigma_next - sigma_hat), ci, cs, co, ci2, cs2, co2, 1. if second else 0., 0., 0., 0.])
        coefs = torch.tensor(rows, dtype=torch.float64).to(torch.float32).to(device).contiguous()
        return coefs, torch.cat(times).to(device), sigmas[0]

    # ---- one cascade stage (elucidated_imagen.py:392-545) ------------------------------------------------
    @torch.no_grad()
    def one_unet_sample(self, unet, shape, *, unet_number, clamp=True, dynamic_threshold=True, cond_scale=1., use_tqdm=True,
                        inpaint_videos=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5, init_images=None,
                        skip_steps=None, sigma_min=None, sigma_max=None, text_embeds=None, text_mask=None, lowres_cond_img=None,
                        lowres_noise_times=None, cond_images=None, **unsupported):
        inpaint_images = default(inpaint_videos, inpaint_images)
        for name, val in unsupported.items():
            if exists(val):
                raise NotImplementedError(f'one_unet_sample({name}=...) is outside the B200 sampling hot path')
        if not clamp:
            raise NotImplementedError('clamp=False')
        assert not (cond_scale != 1. and not self.can_classifier_guidance)
        device = self.device
        require_cuda(device)
        hp = self.hparams[unet_number - 1]
        sigma_min, sigma_max = default(sigma_min, hp.sigma_min), default(sigma_max, hp.sigma_max)
        B, Cimg, H, W = shape
        R = B if cond_scale == 1 else 2 * B
        N = hp.num_sample_steps
        with torch.cuda.device(device):
            coefs, times, init_sigma = self._edm_tables(hp, sigma_min, sigma_max, device)
            plan = unet.plan(R, B, H, W, int(times.numel()), device)
            keep = torch.cat((torch.ones(B, dtype=torch.bool, device=device), torch.zeros(R - B, dtype=torch.bool, device=device)))
            # NOTE: the reference hands the RAW low-res noise level to the U-Net here, not its log-SNR (:700, :727-728)
            plan.prepare(times, text_embeds=text_embeds, text_mask=text_mask, keep=keep, lowres_cond_img=lowres_cond_img,
                         lowres_noise_times=lowres_noise_times, cond_images=cond_images)
            chw = Cimg * H * W
            q_lo, q_hi, q_w = quantile_ranks(chw, self.dynamic_thresholding_percentile, device)
            st_ = plan.sampler_state.setdefault('edm', {})
            if 'x' not in st_:
                for name in ('x', 'x_hat', 'x1', 'd', 'eps'):
                    st_[name] = torch.empty(shape, dtype=torch.float32, device=device)
                st_['coefs'] = torch.empty((4096, 16), dtype=torch.float32, device=device)
                st_['step_ctr'] = torch.zeros(2, dtype=torch.int32, device=device)
                st_['graphs'] = {}
            assert coefs.shape[0] <= 4096
            x, x_hat, x1, d, eps, step_ctr = (st_[n] for n in ('x', 'x_hat', 'x1', 'd', 'eps', 'step_ctr'))
            st_['coefs'][:coefs.shape[0]].copy_(coefs)
            coefs = st_['coefs']
            x.copy_(init_sigma * torch.randn(shape, device=device))            # :442
            if exists(init_images):
                x.add_(init_images.to(device=device, dtype=torch.float32))     # :446-447
            skip = default(skip_steps, 0)                                      # :476-477
            assert 0 <= skip <= N
            step_ctr.fill_(skip)                                               # step index (committed, staged)
            if skip:
                plan.slots.fill_(2 * skip)                                     # two network evaluations per skipped (non-final) step
            has_inpainting = exists(inpaint_images) and exists(inpaint_masks)
            if has_inpainting:                                                 # :455-462
                known = self.resize_to(self.normalize_img(inpaint_images.to(device=device, dtype=torch.float32)), W).contiguous()
                mask = self.resize_to(inpaint_masks.to(device)[:, None].float(), W).bool().to(torch.uint8).contiguous()
                assert known.shape == tuple(shape) and mask.shape == (B, 1, H, W)
            net_in = plan.x_in
            lib = plan.lib
            thr = int(bool(dynamic_threshold))
            key = (float(cond_scale), thr, q_lo, q_hi, q_w)
            sc_ptr = None
            if unet.self_cond:                                                 # latest denoiser output conditions the next evaluation (:496, :518, :538)
                plan.sc_in.zero_()
                sc_ptr = plan.sc_in.data_ptr()

            def phase(ph):
                _lib.check(lib.b200_edm_phase_sc(ph, x.data_ptr(), x_hat.data_ptr(), x1.data_ptr(), d.data_ptr(), net_in.data_ptr(),
                                                 plan.pred.data_ptr(), eps.data_ptr(), coefs.data_ptr(), step_ctr.data_ptr(),
                                                 plan.slots.data_ptr(), R, B, chw, float(cond_scale), thr, q_lo, q_hi, q_w, sc_ptr,
                                                 torch.cuda.current_stream(device).cuda_stream), 'b200_edm_phase')

            def full_step():
                eps.copy_(torch.randn(shape, device=device))                   # :489 (S_noise applied in phase 0)
                phase(0)
                plan.launch()
                phase(1)
                plan.launch()
                phase(2)

            def last_step():                                                   # sigma_next == 0: no second-order correction (:515)
                eps.copy_(torch.randn(shape, device=device))
                phase(0)
                plan.launch()
                phase(1)

            out = torch.empty(shape, dtype=torch.float32, device=device)
            if has_inpainting:
                # RePaint (:481-536): each step is resampled at the same sigma, so the host drives the loop (eager launches) and
                # rewrites the device-side step / slot counters before every resample
                sig = self.sample_schedule(N, hp.rho, sigma_min, sigma_max).tolist()
                st = torch.cuda.current_stream(device).cuda_stream
                launches = 0
                for i in range(skip, N):
                    last_t = i == N - 1
                    for r in reversed(range(inpaint_resample_times)):
                        step_ctr.fill_(i)
                        plan.slots.fill_(2 * i)
                        # x_hat = x + added_noise; masked pixels must become known + added_noise (:498-499): paste `known` into x first
                        _lib.check(lib.b200_inpaint_mix(x.data_ptr(), known.data_ptr(), mask.data_ptr(), None, 1.0, 0.0, B, Cimg, H * W, st),
                                   'b200_inpaint_mix')
                        (last_step if last_t else full_step)()
                        launches += (1 if last_t else 2) * plan.n_launches + 8
                        if not (r == 0 or last_t):                             # :533-536
                            rn = torch.randn(shape, device=device)
                            _lib.check(lib.b200_renoise(x.data_ptr(), rn.data_ptr(), 1.0, sig[i] - sig[i + 1], 1.0, x.numel(), st), 'b200_renoise')
                self.last_launch_count = launches
                _lib.check(lib.b200_finalize_images(x.data_ptr(), x.data_ptr(), x.numel(), 0, st), 'b200_finalize_images')     # clamp :538
                _lib.check(lib.b200_inpaint_mix(x.data_ptr(), known.data_ptr(), mask.data_ptr(), None, 1.0, 0.0, B, Cimg, H * W, st),
                           'b200_inpaint_mix')                                 # :541-542
                _lib.check(lib.b200_finalize_images(x.data_ptr(), out.data_ptr(), out.numel(), 2 | int(self.auto_normalize_img), st),
                           'b200_finalize_images')
                return out
            n_full = max(0, N - 1 - skip)
            it = range(n_full)
            if use_tqdm:
                try:
                    from tqdm.auto import tqdm
                    it = tqdm(it, desc='sampling time step', total=N)
                except ImportError:
                    pass
            if _use_graph() and n_full > 1:
                graph = st_['graphs'].get(key)
                if graph is None:
                    plan.launch()
                    torch.cuda.synchronize(device)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        full_step()
                    st_['graphs'][key] = graph
                for _ in it:
                    graph.replay()
            else:
                for _ in it:
                    full_step()
            if skip < N:
                last_step()
            self.last_launch_count = n_full * (2 * plan.n_launches + 5) + plan.n_launches + 4
            _lib.check(lib.b200_finalize_images(x.data_ptr(), out.data_ptr(), out.numel(), int(self.auto_normalize_img),
                                                torch.cuda.current_stream(device).cuda_stream), 'b200_finalize_images')   # :540-545
        return out

    @torch.no_grad()
    def sample(self, texts=None, text_masks=None, text_embeds=None, cond_images=None, cond_video_frames=None, post_cond_video_frames=None,
               inpaint_videos=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5, init_images=None, skip_steps=None,
               sigma_min=None, sigma_max=None, video_frames=None, batch_size=1, cond_scale=1., lowres_sample_noise_level=None,
               start_at_unet_number=1, start_image_or_video=None, stop_at_unet_number=None, return_all_unet_outputs=False,
               return_pil_images=False, use_tqdm=True, use_one_unet_in_gpu=True, device=None):
        """ElucidatedImagen.sample (elucidated_imagen.py:547-751)."""
        was_training = self.training
        self.eval()
        try:
            device = default(device, self.device)
            self.reset_unets_all_one_device(device=device)
            text_embeds, text_masks = self._encode_texts(texts, text_embeds, text_masks, device)   # elucidated_imagen.py:583-589
            if exists(cond_images) and cond_images.dtype == torch.uint8:      # cast_uint8_images_to_float (elucidated_imagen.py:581)
                cond_images = cond_images / 255
            self._check_cond_images(cond_images, start_at_unet_number, stop_at_unet_number)
            self._check_sample_args(texts, text_embeds, text_masks, dict(cond_video_frames=cond_video_frames,
                                    post_cond_video_frames=post_cond_video_frames, inpaint_videos=inpaint_videos, video_frames=video_frames))
            if return_pil_images:
                raise NotImplementedError('return_pil_images: convert the returned tensor yourself')
            device = next(self.parameters()).device
            if not self.unconditional:
                text_embeds = text_embeds.to(device)
                text_masks = default(text_masks, lambda: torch.any(text_embeds != 0., dim=-1)).to(device)
                batch_size = text_embeds.shape[0]
            lowres_sample_noise_level = default(lowres_sample_noise_level, self.lowres_sample_noise_level)
            num_unets = len(self.unets)
            cond_scale = cast_tuple(cond_scale, num_unets)
            sigma_min, sigma_max = cast_tuple(sigma_min, num_unets), cast_tuple(sigma_max, num_unets)
            if exists(inpaint_images):                                         # elucidated_imagen.py:608-613
                if self.unconditional:
                    batch_size = inpaint_images.shape[0]
                assert inpaint_images.shape[0] == batch_size, 'number of inpainting images must be equal to the specified batch size on sample'
            init_images = cast_tuple(init_images, num_unets)                   # :640-643
            init_images = [self.normalize_img(im) if exists(im) else None for im in init_images]
            skip_steps = cast_tuple(skip_steps, num_unets)
            img = None
            if start_at_unet_number > 1:
                assert start_at_unet_number <= num_unets, 'must start a unet that is less than the total number of unets'
                assert not exists(stop_at_unet_number) or start_at_unet_number <= stop_at_unet_number
                assert exists(start_image_or_video), 'starting image or video must be supplied if only doing upscaling'
                img = self.resize_to(start_image_or_video.to(device), self.image_sizes[start_at_unet_number - 2])
            outputs, launches = [], 0
            for unet_number, unet, image_size, dynamic_threshold, unet_cond_scale, unet_sigma_min, unet_sigma_max, unet_init_images, \
                    unet_skip_steps in zip(range(1, num_unets + 1), self.unets, self.image_sizes, self.dynamic_thresholding, cond_scale, sigma_min,
                                           sigma_max, init_images, skip_steps):
                if unet_number < start_at_unet_number:
                    continue
                assert not isinstance(unet, NullUnet), 'cannot sample from null unet'
                lowres_cond_img = lowres_noise_times = None
                if unet.lowres_cond:
                    lowres_cond_img, lowres_noise_times = self._lowres_conditioning(img, image_size, batch_size, lowres_sample_noise_level, device)
                shape = (batch_size, self.channels, image_size, image_size)
                if exists(unet_init_images):
                    unet_init_images = self.resize_to(unet_init_images.to(device), image_size)   # :709-710
                img = self.one_unet_sample(unet, shape, unet_number=unet_number, text_embeds=text_embeds, text_mask=text_masks,
                                           inpaint_images=inpaint_images, inpaint_masks=inpaint_masks,
                                           inpaint_resample_times=inpaint_resample_times, init_images=unet_init_images,
                                           skip_steps=unet_skip_steps,
                                           sigma_min=unet_sigma_min, sigma_max=unet_sigma_max, cond_scale=unet_cond_scale,
                                           lowres_cond_img=lowres_cond_img, lowres_noise_times=lowres_noise_times,
                                           dynamic_threshold=dynamic_threshold, use_tqdm=use_tqdm, cond_images=cond_images)
                launches += self.last_launch_count
                outputs.append(img)
                if exists(stop_at_unet_number) and stop_at_unet_number == unet_number:
                    break
            self.last_launch_count = launches
            return outputs if return_all_unet_outputs else outputs[-1]
        finally:
            self.train(was_training)
