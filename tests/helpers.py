"""Shared test helpers: golden fixture loading and oracle-side model reconstruction."""
import os

import torch

from oracle import unet_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def contract():
    return load_golden('state_dict_contract.pt')


def synth_weights(contract_name, seed):
    return unet_ref.synth_state_dict(contract()[contract_name], seed=seed)


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp(min=1e-12)).item()


def seeded_inputs(seed, B, size, L=256, D=768, ragged=True):
    """Inputs of the BASELINE-shape fixtures (tests/golden/baseline_shapes.pt); identical to oracle/make_golden.py's helper:
    a CPU torch.Generator is bit-reproducible across machines, so only the reference OUTPUTS are stored."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, size, size, generator=g)
    te = torch.randn(B, L, D, generator=g)
    if ragged and B > 1:
        te[1, L // 3:] = 0.
    return x, te, torch.any(te != 0., dim=-1)


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}
