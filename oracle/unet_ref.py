"""ORACLE (test infrastructure, NOT product code) -- CPU fp32 restatement of the
reference's U-Net forward for the sampling hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product path (imagen_pytorch_b200) never does.

The arithmetic is plain torch fp32 because that *is* the reference's arithmetic
(the reference is pure PyTorch).  Every function cites the
reference lines it restates (paths relative to /root/reference/imagen_pytorch/).
It is functional: it consumes a reference-layout ``state_dict`` plus the Unet
constructor kwargs, so it can be pinned against the real reference (see
oracle/make_golden.py, which imports /root/reference in the build container and
writes tests/golden/*.pt) and can travel to the GPU box without the reference.

Parity status: PINNED against the live reference (bit-level fp32 agreement on CPU
checked by oracle/make_golden.py at generation time; fixtures committed).
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- config


def unet_config(**kw):
    """Defaults of Unet.__init__ (imagen_pytorch.py:1113-1161)."""
    cfg = dict(
        dim=None, text_embed_dim=768, num_resnet_blocks=1, cond_dim=None,
        num_time_tokens=2, learned_sinu_pos_emb_dim=16, dim_mults=(1, 2, 4, 8),
        cond_images_channels=0, channels=3, channels_out=None, attn_dim_head=64,
        attn_heads=8, ff_mult=2., lowres_cond=False, layer_attns=True,
        layer_attns_depth=1, layer_mid_attns_depth=1, attend_at_middle=True,
        layer_cross_attns=True, cond_on_text=True, max_text_len=256, init_dim=None,
        init_cross_embed=True, init_cross_embed_kernel_sizes=(3, 7, 15),
        init_conv_kernel_size=7, attn_pool_text=True, attn_pool_num_latents=32,
        memory_efficient=False, use_global_context_attn=True,
        scale_skip_connection=True, final_resnet_block=True,
        final_conv_kernel_size=3, pixel_shuffle_upsample=True, self_cond=False,
    )
    for k, v in kw.items():
        cfg[k] = v
    assert cfg['dim'] is not None
    return cfg


def _tup(v, n):
    if isinstance(v, list):
        v = tuple(v)
    return v if isinstance(v, tuple) else (v,) * n


# --------------------------------------------------------------------------- pieces


def _l2norm(t):                      # imagen_pytorch.py:133-134
    return F.normalize(t, dim=-1)


def _chan_rmsnorm(x, gamma):         # ChanRMSNorm, imagen_pytorch.py:322-329
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def _ln_gain(x, g):                  # custom LayerNorm (gain only), imagen_pytorch.py:331-349
    var = torch.var(x, dim=-1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=-1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _nn_ln(x, sd, p):                # nn.LayerNorm (affine, eps 1e-5)
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _linear(x, sd, p):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def _feedforward(x, sd, p):          # FeedForward, imagen_pytorch.py:972-980
    h = _ln_gain(x, sd[p + '.0.g'])
    h = F.linear(h, sd[p + '.1.weight'])
    h = F.gelu(h)
    h = _ln_gain(h, sd[p + '.3.g'])
    return F.linear(h, sd[p + '.4.weight'])


def _split_heads(t, h):
    b, n, hd = t.shape
    return t.view(b, n, h, hd // h).permute(0, 2, 1, 3)


def _merge_heads(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def _perceiver_attention(x, latents, sd, p, heads):   # imagen_pytorch.py:408-445
    x = _nn_ln(x, sd, p + '.norm')
    latents = _nn_ln(latents, sd, p + '.norm_latents')
    q = F.linear(latents, sd[p + '.to_q.weight'])
    kv = F.linear(torch.cat((x, latents), dim=-2), sd[p + '.to_kv.weight'])
    k, v = kv.chunk(2, dim=-1)
    q, k, v = (_split_heads(t, heads) for t in (q, k, v))
    q = _l2norm(q) * sd[p + '.q_scale']
    k = _l2norm(k) * sd[p + '.k_scale']
    sim = torch.einsum('bhid,bhjd->bhij', q, k) * 8
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    out = _merge_heads(torch.einsum('bhij,bhjd->bhid', attn, v))
    out = F.linear(out, sd[p + '.to_out.0.weight'])
    return _nn_ln(out, sd, p + '.to_out.1')


def _perceiver_resampler(x, sd, p, heads):            # imagen_pytorch.py:481-498
    n = x.shape[1]
    x_pos = x + sd[p + '.pos_emb.weight'][:n]
    latents = sd[p + '.latents'].unsqueeze(0).expand(x.shape[0], -1, -1)
    mp = p + '.to_latents_from_mean_pooled_seq'
    if (mp + '.1.weight') in sd:
        pooled = x.sum(dim=1) / torch.full((x.shape[0], 1), float(n), device=x.device).clamp(min=1e-5)   # masked_mean with all-ones mask :490,:142-150
        ml = _ln_gain(pooled, sd[mp + '.0.g'])
        ml = _linear(ml, sd, mp + '.1')
        ml = ml.view(x.shape[0], -1, x.shape[-1])
        latents = torch.cat((ml, latents), dim=-2)
    layer = 0
    while f'{p}.layers.{layer}.0.to_q.weight' in sd:
        lp = f'{p}.layers.{layer}'
        latents = _perceiver_attention(x_pos, latents, sd, lp + '.0', heads) + latents
        latents = _feedforward(latents, sd, lp + '.1') + latents
        layer += 1
    return latents


def _self_attention(x, context, sd, p, heads):        # Attention.forward, imagen_pytorch.py:534-591
    b, n, _ = x.shape
    x = _ln_gain(x, sd[p + '.norm.g'])
    q = F.linear(x, sd[p + '.to_q.weight'])
    k, v = F.linear(x, sd[p + '.to_kv.weight']).chunk(2, dim=-1)
    q = _split_heads(q, heads)
    nk, nv = sd[p + '.null_kv'].unbind(dim=-2)
    k = torch.cat((nk.expand(b, 1, -1), k), dim=-2)
    v = torch.cat((nv.expand(b, 1, -1), v), dim=-2)
    if context is not None:
        c = _nn_ln(context, sd, p + '.to_context.0')
        ck, cv = _linear(c, sd, p + '.to_context.1').chunk(2, dim=-1)
        k = torch.cat((ck, k), dim=-2)
        v = torch.cat((cv, v), dim=-2)
    q = _l2norm(q) * sd[p + '.q_scale']
    k = _l2norm(k) * sd[p + '.k_scale']
    sim = torch.einsum('bhid,bjd->bhij', q, k) * 8
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    out = _merge_heads(torch.einsum('bhij,bjd->bhid', attn, v))
    out = F.linear(out, sd[p + '.to_out.0.weight'])
    return _ln_gain(out, sd[p + '.to_out.1.g'])


def _cross_attention(x, context, sd, p, heads):       # CrossAttention.forward, imagen_pytorch.py:793-834
    b = x.shape[0]
    x = _ln_gain(x, sd[p + '.norm.g'])
    q = F.linear(x, sd[p + '.to_q.weight'])
    k, v = F.linear(context, sd[p + '.to_kv.weight']).chunk(2, dim=-1)
    q, k, v = (_split_heads(t, heads) for t in (q, k, v))
    nk, nv = sd[p + '.null_kv'].unbind(dim=-2)
    k = torch.cat((nk.expand(b, heads, 1, -1), k), dim=-2)
    v = torch.cat((nv.expand(b, heads, 1, -1), v), dim=-2)
    q = _l2norm(q) * sd[p + '.q_scale']
    k = _l2norm(k) * sd[p + '.k_scale']
    sim = torch.einsum('bhid,bhjd->bhij', q, k) * 8
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    out = _merge_heads(torch.einsum('bhij,bhjd->bhid', attn, v))
    out = F.linear(out, sd[p + '.to_out.0.weight'])
    return _ln_gain(out, sd[p + '.to_out.1.g'])


def _transformer_block(x, context, sd, p, heads):     # TransformerBlock.forward, imagen_pytorch.py:1012-1022
    b, c, h, w = x.shape
    t = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    layer = 0
    while f'{p}.layers.{layer}.0.to_q.weight' in sd:
        lp = f'{p}.layers.{layer}'
        t = _self_attention(t, context, sd, lp + '.0', heads) + t
        t = _feedforward(t, sd, lp + '.1') + t
        layer += 1
    return t.view(b, h, w, c).permute(0, 3, 1, 2)


def _block(x, sd, p, scale_shift=None):               # Block.forward, imagen_pytorch.py:683-691
    x = _chan_rmsnorm(x, sd[p + '.norm.gamma'])
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    x = F.silu(x)
    return F.conv2d(x, sd[p + '.project.weight'], sd[p + '.project.bias'], padding=1)


def _global_context(x, sd, p):                        # GlobalContext.forward, imagen_pytorch.py:965-970
    context = F.conv2d(x, sd[p + '.to_k.weight'], sd[p + '.to_k.bias'])
    b, c = x.shape[:2]
    xf, cf = x.reshape(b, c, -1), context.reshape(b, 1, -1)
    out = torch.einsum('bin,bcn->bci', cf.softmax(dim=-1), xf).unsqueeze(-1)
    out = F.conv2d(out, sd[p + '.net.0.weight'], sd[p + '.net.0.bias'])
    out = F.silu(out)
    out = F.conv2d(out, sd[p + '.net.2.weight'], sd[p + '.net.2.bias'])
    return out.sigmoid()


def _resnet_block(x, sd, p, heads, time_emb=None, cond=None):   # ResnetBlock.forward, imagen_pytorch.py:735-757
    scale_shift = None
    if (p + '.time_mlp.1.weight') in sd and time_emb is not None:
        te = _linear(F.silu(time_emb), sd, p + '.time_mlp.1')
        scale_shift = te[:, :, None, None].chunk(2, dim=1)
    h = _block(x, sd, p + '.block1')
    if (p + '.cross_attn.to_q.weight') in sd:
        assert cond is not None
        b, c, hh, ww = h.shape
        t = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        t = _cross_attention(t, cond, sd, p + '.cross_attn', heads) + t
        h = t.view(b, hh, ww, c).permute(0, 3, 1, 2)
    h = _block(h, sd, p + '.block2', scale_shift)
    if (p + '.gca.to_k.weight') in sd:
        h = h * _global_context(h, sd, p + '.gca')
    if (p + '.res_conv.weight') in sd:
        x = F.conv2d(x, sd[p + '.res_conv.weight'], sd[p + '.res_conv.bias'])
    return h + x


def _sinu_pos_emb(x, weights):                        # LearnedSinusoidalPosEmb.forward, imagen_pytorch.py:664-669
    x = x[:, None]
    freqs = x * weights[None, :] * 2 * math.pi
    return torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)


def _downsample(x, sd, p):                            # Downsample, imagen_pytorch.py:633-640
    b, c, h, w = x.shape
    x = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, h // 2, w // 2)
    return F.conv2d(x, sd[p + '.1.weight'], sd[p + '.1.bias'])


def _upsample(x, sd, p, pixel_shuffle):               # PixelShuffleUpsample :603-631 / Upsample :595-601
    if pixel_shuffle:
        x = F.conv2d(x, sd[p + '.net.0.weight'], sd[p + '.net.0.bias'])
        return F.pixel_shuffle(F.silu(x), 2)
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    return F.conv2d(x, sd[p + '.1.weight'], sd[p + '.1.bias'], padding=1)


# --------------------------------------------------------------------------- forward


def unet_forward(sd, cfg, x, time, *, text_embeds=None, text_mask=None,
                 lowres_cond_img=None, lowres_noise_times=None, cond_drop_prob=0., self_cond=None, cond_images=None):
    """Unet.forward, imagen_pytorch.py:1524-1725 (combine_upsample_fmaps / init_conv_to_final_conv_residual
    branches omitted: out of scope per SURVEY.md section 8a; the product rejects them too)."""
    heads = cfg['attn_heads']
    batch = x.shape[0]
    nlev = len(cfg['dim_mults'])
    nrb = _tup(cfg['num_resnet_blocks'], nlev)
    layer_attns = _tup(cfg['layer_attns'], nlev)
    mem_eff = cfg['memory_efficient']

    if cfg.get('self_cond', False):                                            # :1541-1543
        x = torch.cat((x, torch.zeros_like(x) if self_cond is None else self_cond), dim=1)
    assert not (cfg['lowres_cond'] and lowres_cond_img is None)
    if lowres_cond_img is not None:
        x = torch.cat((x, lowres_cond_img), dim=1)                             # :1550-1551
    assert not ((cfg.get('cond_images_channels', 0) > 0) ^ (cond_images is not None))           # :1555
    if cond_images is not None:                                                # :1557-1560
        assert cond_images.shape[1] == cfg['cond_images_channels']
        if cond_images.shape[-1] != x.shape[-1]:
            cond_images = F.interpolate(cond_images, x.shape[-1], mode=cfg.get('resize_mode', 'nearest'))   # resize_image_to :152-168
        x = torch.cat((cond_images, x), dim=1)

    if cfg['init_cross_embed']:                                                # CrossEmbedLayer :1074-1076
        ks = sorted(cfg['init_cross_embed_kernel_sizes'])
        x = torch.cat([F.conv2d(x, sd[f'init_conv.convs.{i}.weight'], sd[f'init_conv.convs.{i}.bias'],
                                padding=(k - 1) // 2) for i, k in enumerate(ks)], dim=1)
    else:
        k = cfg['init_conv_kernel_size']
        x = F.conv2d(x, sd['init_conv.weight'], sd['init_conv.bias'], padding=k // 2)

    # time conditioning :1573-1589
    time_hiddens = F.silu(_linear(_sinu_pos_emb(time, sd['to_time_hiddens.0.weights']), sd, 'to_time_hiddens.1'))
    time_tokens = _linear(time_hiddens, sd, 'to_time_tokens.0').view(batch, cfg['num_time_tokens'], -1)
    t = _linear(time_hiddens, sd, 'to_time_cond.0')
    if cfg['lowres_cond']:
        lh = F.silu(_linear(_sinu_pos_emb(lowres_noise_times, sd['to_lowres_time_hiddens.0.weights']), sd, 'to_lowres_time_hiddens.1'))
        ltok = _linear(lh, sd, 'to_lowres_time_tokens.0').view(batch, cfg['num_time_tokens'], -1)
        t = t + _linear(lh, sd, 'to_lowres_time_cond.0')
        time_tokens = torch.cat((time_tokens, ltok), dim=-2)

    # text conditioning :1595-1652
    text_tokens = None
    if text_embeds is not None and cfg['cond_on_text']:
        keep = torch.full((batch,), cond_drop_prob != 1., dtype=torch.bool, device=x.device)    # prob_mask_like :201-207 (prob in {0,1})
        assert cond_drop_prob in (0., 1.), 'oracle covers the sampling path only (no random dropout)'
        max_len = cfg['max_text_len']
        text_tokens = _linear(text_embeds, sd, 'text_to_cond')[:, :max_len]
        keep_embed = keep[:, None, None]
        if text_mask is not None:
            text_mask = text_mask[:, :max_len]
        rem = max_len - text_tokens.shape[1]
        if rem > 0:
            text_tokens = F.pad(text_tokens, (0, 0, 0, rem))
        if text_mask is not None:
            if rem > 0:
                text_mask = F.pad(text_mask, (0, rem), value=False)
            keep_embed = text_mask[:, :, None] & keep_embed
        text_tokens = torch.where(keep_embed, text_tokens, sd['null_text_embed'])
        if cfg['attn_pool_text']:
            text_tokens = _perceiver_resampler(text_tokens, sd, 'attn_pool', heads)
        th = text_tokens.mean(dim=-2)
        th = _nn_ln(th, sd, 'to_text_non_attn_cond.0')
        th = _linear(F.silu(_linear(th, sd, 'to_text_non_attn_cond.1')), sd, 'to_text_non_attn_cond.3')
        th = torch.where(keep[:, None], th, sd['null_text_hidden'])
        t = t + th

    c = time_tokens if text_tokens is None else torch.cat((time_tokens, text_tokens), dim=-2)
    c = _nn_ln(c, sd, 'norm_cond')                                             # :1656-1660

    if mem_eff:
        x = _resnet_block(x, sd, 'init_resnet_block', heads, t)                # :1664-1665

    hiddens = []
    for i in range(nlev):                                                      # :1671-1685
        p = f'downs.{i}'
        if mem_eff:
            x = _downsample(x, sd, p + '.0')
        x = _resnet_block(x, sd, p + '.1', heads, t, c)
        for j in range(nrb[i]):
            x = _resnet_block(x, sd, f'{p}.2.{j}', heads, t)
            hiddens.append(x)
        if layer_attns[i]:
            x = _transformer_block(x, c, sd, p + '.3', heads)
        hiddens.append(x)
        if not mem_eff:
            if i < nlev - 1:
                x = _downsample(x, sd, p + '.4')
            else:                                                              # Parallel(conv3x3, conv1x1) :1366, :368-375
                x = F.conv2d(x, sd[p + '.4.fns.0.weight'], sd[p + '.4.fns.0.bias'], padding=1) + \
                    F.conv2d(x, sd[p + '.4.fns.1.weight'], sd[p + '.4.fns.1.bias'])

    x = _resnet_block(x, sd, 'mid_block1', 8 if 'mid_block1.cross_attn.to_q.weight' in sd else heads, t, c)   # heads default 8 (:1380 uses ResnetBlock, not resnet_klass)
    if cfg['attend_at_middle']:
        x = _transformer_block(x, None, sd, 'mid_attn', heads)                 # :1689-1690
    x = _resnet_block(x, sd, 'mid_block2', 8 if 'mid_block2.cross_attn.to_q.weight' in sd else heads, t, c)

    skip_scale = 2 ** -0.5 if cfg['scale_skip_connection'] else 1.
    for i in range(nlev):                                                      # :1698-1708
        p = f'ups.{i}'
        lev = nlev - 1 - i
        x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
        x = _resnet_block(x, sd, p + '.0', heads, t, c)
        for j in range(nrb[lev]):
            x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
            x = _resnet_block(x, sd, f'{p}.1.{j}', heads, t)
        if layer_attns[lev]:
            x = _transformer_block(x, c, sd, p + '.2', heads)
        if (i < nlev - 1) or mem_eff:
            x = _upsample(x, sd, p + '.3', cfg['pixel_shuffle_upsample'])

    if cfg['final_resnet_block']:
        x = _resnet_block(x, sd, 'final_res_block', heads, t)                  # :1719-1720
    if lowres_cond_img is not None:
        x = torch.cat((x, lowres_cond_img), dim=1)                             # :1722-1723
    k = cfg['final_conv_kernel_size']
    return F.conv2d(x, sd['final_conv.weight'], sd['final_conv.bias'], padding=k // 2)


def unet_forward_with_cond_scale(sd, cfg, x, time, *, cond_scale=1., **kw):
    """Unet.forward_with_cond_scale, imagen_pytorch.py:1510-1522."""
    logits = unet_forward(sd, cfg, x, time, **kw)
    if cond_scale == 1:
        return logits
    null_logits = unet_forward(sd, cfg, x, time, cond_drop_prob=1., **kw)
    return null_logits + (logits - null_logits) * cond_scale


def synth_state_dict(shapes, seed=0):
    """Deterministic, non-degenerate weights for a reference-layout state_dict.

    The reference zero-initialises final_conv (imagen_pytorch.py:1438) so a freshly
    constructed U-Net outputs exactly 0 and every norm gain is exactly 1: parity on
    fresh weights would be vacuous.  Instead of shipping ~50 MB of
    reference-initialised weights per fixture, goldens are generated from weights
    synthesised here from the (key -> shape) contract alone; the same call rebuilds
    them bit-identically on the GPU box (CPU torch.Generator)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        r = torch.randn(shp, generator=g)
        leaf = k.rsplit('.', 1)[-1]
        if leaf in ('gamma', 'g', 'q_scale', 'k_scale') or (leaf == 'weight' and len(shp) == 1):
            v = 1 + 0.2 * r                               # norm gains / qk scales
        elif leaf == 'bias':
            v = 0.05 * r
        elif leaf in ('weights', 'null_kv', 'latents', 'null_text_embed', 'null_text_hidden') or 'pos_emb' in k:
            v = r
        elif leaf == 'weight':
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            v = r / math.sqrt(fan_in)
        else:
            raise KeyError(f'unclassified parameter {k}')
        sd[k] = v
    return sd
