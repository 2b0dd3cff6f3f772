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
