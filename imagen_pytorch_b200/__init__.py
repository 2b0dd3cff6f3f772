"""imagen_pytorch_b200 -- B200-native drop-in for the sampling hot path of lucidrains/imagen-pytorch.

Public names mirror the reference package (imagen_pytorch/__init__.py) for the path in scope:
``Unet``, ``BaseUnet64``, ``SRUnet256``, ``SRUnet1024``, ``NullUnet``, ``Imagen``, ``ElucidatedImagen``.
"""
from .unet import Unet, NullUnet, BaseUnet64, SRUnet256, SRUnet1024, UnetPlan
from .imagen import Imagen, GaussianDiffusionContinuousTimes
from .elucidated import ElucidatedImagen
from .dist import sample_sharded, sample_in_chunks
from .trainer import TrainedSampler, split_trainer_checkpoint
from . import t5
from ._lib import B200Error, LIB_PATH

__version__ = '0.1.0'
