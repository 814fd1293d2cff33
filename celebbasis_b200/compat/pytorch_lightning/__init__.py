"""Minimal in-repo stand-in for the `pytorch_lightning==1.5.9` surface the reference drivers touch (SURVEY.md 8b / 8f-1:
main_id_embed.py:7-19,183,217,295,344,450,540,640,658,748,812; scripts/stable_txt2img.py:11; ldm/models/diffusion/
ddpm.py:14,21; ldm/models/autoencoder.py:2).  One process per GPU (launched by torchrun), the optimisation loop is the
plain `training_step -> backward -> all-reduce(mean) of the trainable grads -> optimizer.step` the reference gets from
Lightning DDP.  Used only when the real package is not installed (celebbasis_b200.compat.install())."""
import os
import random

import numpy as np
import torch

from .core import Callback, LightningDataModule, LightningModule
from .trainer import Trainer
from . import callbacks, loggers, utilities  # noqa: F401

__version__ = "1.5.9"


def seed_everything(seed=None, workers=False):
    seed = int(seed if seed is not None else os.environ.get("PL_GLOBAL_SEED", 0))
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


__all__ = ["Callback", "LightningDataModule", "LightningModule", "Trainer", "seed_everything", "callbacks", "loggers",
           "utilities"]
