"""LightningModule / LightningDataModule / Callback protocol stand-ins."""
import torch
from torch import nn


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.trainer = None
        self._global_step = 0
        self._current_epoch = 0
        self.last_log = {}

    # -- state Lightning exposes as read-only properties; plain attributes here so tests / drivers may set them
    @property
    def global_step(self):
        return self.trainer.global_step if self.trainer is not None else self._global_step

    @global_step.setter
    def global_step(self, v):
        self._global_step = int(v)

    @property
    def current_epoch(self):
        return self.trainer.current_epoch if self.trainer is not None else self._current_epoch

    @current_epoch.setter
    def current_epoch(self, v):
        self._current_epoch = int(v)

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    @property
    def global_rank(self):
        return self.trainer.global_rank if self.trainer is not None else 0

    @property
    def logger(self):
        return self.trainer.logger if self.trainer is not None else None

    def log(self, name, value, *args, **kwargs):
        self.last_log[name] = value.detach() if torch.is_tensor(value) else value
        if self.trainer is not None:
            self.trainer.logged_metrics[name] = self.last_log[name]

    def log_dict(self, d, *args, **kwargs):
        for k, v in dict(d).items():
            self.log(k, v)

    def optimizers(self):
        opts = self.trainer.optimizers if self.trainer is not None else []
        return opts[0] if len(opts) == 1 else opts

    # -- hooks (no-ops by default)
    def on_train_batch_start(self, batch, batch_idx, dataloader_idx=0):
        pass

    def on_train_batch_end(self, outputs, batch, batch_idx, dataloader_idx=0):
        pass

    def on_save_checkpoint(self, checkpoint):
        pass

    def on_load_checkpoint(self, checkpoint):
        pass

    def configure_optimizers(self):
        raise NotImplementedError

    def training_step(self, batch, batch_idx):
        raise NotImplementedError


class LightningDataModule:
    def __init__(self, *args, **kwargs):
        self.trainer = None

    def prepare_data(self):
        pass

    def setup(self, stage=None):
        pass


class Callback:
    """Hook names of Lightning 1.5 used by the reference's callbacks (main_id_embed.py:295-490)."""

    def setup(self, trainer, pl_module, stage=None): pass
    def on_pretrain_routine_start(self, trainer, pl_module): pass
    def on_train_start(self, trainer, pl_module): pass
    def on_train_epoch_start(self, trainer, pl_module): pass
    def on_train_batch_start(self, trainer, pl_module, batch, batch_idx, dataloader_idx=0): pass
    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0): pass
    def on_train_epoch_end(self, trainer, pl_module, *args): pass
    def on_train_end(self, trainer, pl_module): pass
    def on_validation_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0): pass
    def on_keyboard_interrupt(self, trainer, pl_module): pass
    def on_save_checkpoint(self, trainer, pl_module, checkpoint): pass
