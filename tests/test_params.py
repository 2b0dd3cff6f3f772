"""CPU: the generated parameter layout equals the reference's state_dict contract (keys AND shapes)."""
import pytest
import torch

from imagen_pytorch_b200 import Unet, SRUnet256, Imagen, ElucidatedImagen, B200Error
from imagen_pytorch_b200.params import UnetArch, param_table
from tests.helpers import contract, synth_weights

CASES = {
    'base_dim128': dict(dim=128),
    'base_dim32': dict(dim=32, dim_mults=(1, 2, 4, 8)),
    'base_dim192': dict(dim=192),
    'test_dh32_cond': dict(dim=32, dim_mults=(1, 2, 4), text_embed_dim=64, max_text_len=24, attn_dim_head=32, attn_heads=4, cond_images_channels=3),
    'srunet1024_t64': dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), layer_attns=False, layer_cross_attns=(False, False, False, True),
                           attn_heads=8, ff_mult=2., memory_efficient=True, lowres_cond=True, text_embed_dim=64),
    'test_base': dict(dim=32, dim_mults=(1, 2, 4, 8), text_embed_dim=64, max_text_len=24),
    'test_sr': dict(dim=32, dim_mults=(1, 2, 4), text_embed_dim=64, max_text_len=24, num_resnet_blocks=(1, 2, 2),
                    layer_attns=(False, False, True), layer_cross_attns=(False, False, True), memory_efficient=True, lowres_cond=True),
    'test_selfcond': dict(dim=32, dim_mults=(1, 2, 4, 8), text_embed_dim=64, max_text_len=24, self_cond=True),
}


@pytest.mark.parametrize('name', list(CASES))
def test_param_table_matches_reference_contract(name):
    mine = {k: tuple(v[0]) for k, v in param_table(UnetArch(**CASES[name])).items()}
    ref = {k: tuple(v) for k, v in contract()[name].items()}
    assert mine == ref


def test_srunet256_preset_matches_reference_contract():
    u = SRUnet256(lowres_cond=True)
    assert {k: tuple(v.shape) for k, v in u.state_dict().items()} == {k: tuple(v) for k, v in contract()['srunet256'].items()}


def test_reference_state_dict_loads_strict():
    u = Unet(**CASES['test_base'])
    sd = synth_weights('test_base', 0)
    u.load_state_dict(sd, strict=True)
    got = u.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    assert u._locals['dim'] == 32 and u.lowres_cond is False and u.channels == 3


def test_fresh_unet_final_conv_is_zero_like_reference():
    u = Unet(dim=32, dim_mults=(1, 2))
    assert u.final_conv.weight.abs().max() == 0 and u.final_conv.bias.abs().max() == 0   # zero_init_ :1438


@pytest.mark.parametrize('kw', [dict(use_linear_attn=True), dict(init_conv_to_final_conv_residual=True), dict(attn_dim_head=128),
                                dict(pixel_shuffle_upsample=False), dict(combine_upsample_fmaps=True), dict(cross_embed_downsample=True)])
def test_unsupported_options_raise_instead_of_diverging(kw):
    with pytest.raises(NotImplementedError):
        Unet(dim=32, **kw)


def test_cast_model_parameters_reinstantiates_for_cascade():
    u = Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64)
    im = Imagen((u, Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64)), image_sizes=(16, 32), text_embed_dim=64, timesteps=4)
    assert im.unets[0] is u and im.unets[1].lowres_cond is True
    assert 'to_lowres_time_cond.0.weight' in im.unets[1].state_dict()
    assert im.noise_schedulers[0].num_timesteps == 4 and len(im.noise_schedulers) == 2
    keys = set(im.state_dict())
    assert 'unets.0.final_conv.weight' in keys and not any('_temp' in k for k in keys)


def test_cpu_module_fails_loudly_without_fallback():
    u = Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64, max_text_len=8)
    with pytest.raises(B200Error):
        u(torch.randn(1, 3, 16, 16), torch.zeros(1), text_embeds=torch.randn(1, 8, 64))
    with pytest.raises(B200Error):
        Imagen(u, image_sizes=16, text_embed_dim=64, timesteps=2).sample(text_embeds=torch.randn(1, 8, 64), use_tqdm=False)
    with pytest.raises(B200Error):
        ElucidatedImagen(u, image_sizes=16, text_embed_dim=64, num_sample_steps=2).sample(text_embeds=torch.randn(1, 8, 64), use_tqdm=False)


def test_training_forward_is_rejected():
    u = Unet(dim=32, dim_mults=(1, 2), text_embed_dim=64)
    with pytest.raises(NotImplementedError):
        Imagen(u, image_sizes=16, text_embed_dim=64)(torch.randn(1, 3, 16, 16), text_embeds=torch.randn(1, 8, 64))
