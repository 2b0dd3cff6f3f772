"""Trainer-side caller of the sampling hot path.

Training itself is out of scope (train with the reference); what a user of ``ImagenTrainer`` needs in order to SAMPLE on B200 is
  * loading an ``ImagenTrainer.save`` checkpoint (trainer.py:677-736): ``{'model': imagen.state_dict(), 'ema': ema_unets.state_dict(),
    'version', 'steps', ...}`` where the EMA copies live under ``'<unet index>.ema_model.<unet key>'`` (ema_pytorch.EMA inside an
    nn.ModuleList),
  * ``ImagenTrainer.sample`` semantics (trainer.py:947-961): sample with the EMA U-Nets swapped in unless ``use_non_ema=True``
    (``use_ema_unets``, :846-869), and split large batches with ``max_batch_size`` (``imagen_sample_in_chunks``, :188-206).

``TrainedSampler`` provides exactly that around a B200 ``Imagen`` / ``ElucidatedImagen``; every U-Net (online and EMA) stays
resident on the device (180 GB HBM) and keeps its own launch plans, so switching between them costs nothing.
"""
from __future__ import annotations

import copy
from contextlib import contextmanager

import torch
from torch import nn

from .dist import sample_in_chunks
from .unet import Unet


def split_trainer_checkpoint(obj):
    """-> (imagen state_dict, [per-unet EMA state_dict or None]) from a loaded ``ImagenTrainer.save`` object."""
    model = obj['model']
    ema = obj.get('ema')
    n = 1 + max(int(k.split('.')[1]) for k in model if k.startswith('unets.'))
    ema_sds = [None] * n
    if ema is not None:
        for i in range(n):
            pre = f'{i}.ema_model.'
            sd = {k[len(pre):]: v for k, v in ema.items() if k.startswith(pre)}
            ema_sds[i] = sd or None
    return model, ema_sds


class TrainedSampler(nn.Module):
    def __init__(self, imagen, use_ema=True):
        super().__init__()
        self.imagen = imagen
        self.use_ema = use_ema
        self.ema_unets = nn.ModuleList([self._clone(u) for u in imagen.unets]) if use_ema else None

    @staticmethod
    def _clone(unet):
        """A second U-Net of the same architecture and weights (what ema_pytorch.EMA's deepcopy gives the reference); it compiles
        its own launch plans on first use."""
        if not isinstance(unet, Unet):
            return copy.deepcopy(unet)
        twin = Unet(**unet._locals)
        twin.load_state_dict(unet.state_dict())
        return twin.to(next(unet.parameters()).device)

    # ---- ImagenTrainer.load (model / EMA weights only: optimizers, schedulers and scalers are training state)
    def load(self, path_or_obj, strict=True):
        obj = torch.load(path_or_obj, map_location='cpu', weights_only=False) if not isinstance(path_or_obj, dict) else path_or_obj
        model, ema_sds = split_trainer_checkpoint(obj)
        missing, unexpected = self.imagen.load_state_dict(model, strict=False)
        missing = [k for k in missing if k.startswith('unets.')]
        unexpected = [k for k in unexpected if k.startswith('unets.')]
        if strict and (missing or unexpected):
            raise RuntimeError(f'checkpoint does not match the U-Nets: missing {missing[:5]}, unexpected {unexpected[:5]}')
        if self.use_ema:
            for unet, sd in zip(self.ema_unets, ema_sds):
                if sd is None:
                    raise RuntimeError("checkpoint has no 'ema' weights: construct TrainedSampler(..., use_ema=False) or sample with use_non_ema=True")
                unet.load_state_dict(sd, strict=strict)
        return obj.get('steps'), obj.get('version')

    @contextmanager
    def use_ema_unets(self):                                                   # trainer.py:846-869
        if not self.use_ema:
            yield
            return
        self.ema_unets.to(self.imagen.device)
        self.ema_unets.eval()
        online = self.imagen.unets
        self.imagen.unets = self.ema_unets
        try:
            yield
        finally:
            self.imagen.unets = online

    @torch.no_grad()
    def sample(self, *args, max_batch_size=None, use_non_ema=False, **kwargs):  # trainer.py:947-961 (+ :188-206)
        if use_non_ema or not self.use_ema:
            return sample_in_chunks(self.imagen, *args, max_batch_size=max_batch_size, **kwargs)
        with self.use_ema_unets():
            return sample_in_chunks(self.imagen, *args, max_batch_size=max_batch_size, **kwargs)
