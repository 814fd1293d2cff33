"""ModelCheckpoint / LearningRateMonitor stand-ins (main_id_embed.py:17,658-739)."""
import os

from ..core import Callback


class ModelCheckpoint(Callback):
    def __init__(self, dirpath=None, filename="{epoch:06}", verbose=False, save_last=None, save_top_k=1, monitor=None,
                 every_n_train_steps=None, every_n_epochs=None, save_weights_only=False, **kwargs):
        self.dirpath, self.filename, self.verbose = dirpath, filename, verbose
        self.save_last, self.save_top_k, self.monitor = save_last, save_top_k, monitor
        self.every_n_train_steps, self.every_n_epochs = every_n_train_steps, every_n_epochs
        self.last_model_path = ""
        self._last_saved_step = -1

    def _path(self, trainer):
        name = self.filename or "{epoch:06}"
        try:
            name = name.format(epoch=trainer.current_epoch, step=trainer.global_step)
        except (KeyError, IndexError, ValueError):
            name = f"epoch={trainer.current_epoch:06d}"
        return os.path.join(self.dirpath or trainer.default_root_dir, name + ".ckpt")

    def _save(self, trainer, path):
        if trainer.global_rank == 0:
            os.makedirs(os.path.dirname(path), exist_ok=True)
        trainer.save_checkpoint(path)
        self.last_model_path = path
        if self.verbose and trainer.global_rank == 0:
            print(f"[ModelCheckpoint] saved {path}")

    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx, dataloader_idx=0):
        n = self.every_n_train_steps
        if n and trainer.global_step > 0 and trainer.global_step % n == 0 and trainer.global_step != self._last_saved_step:
            self._last_saved_step = trainer.global_step      # accumulate_grad_batches > 1: one save per optimiser step
            self._save(trainer, self._path(trainer))

    def on_train_end(self, trainer, pl_module):
        if self.save_last:
            self._save(trainer, os.path.join(self.dirpath or trainer.default_root_dir, "last.ckpt"))


class LearningRateMonitor(Callback):
    def __init__(self, logging_interval=None, log_momentum=False):
        self.logging_interval = logging_interval

    def on_train_batch_start(self, trainer, pl_module, batch, batch_idx, dataloader_idx=0):
        for i, opt in enumerate(trainer.optimizers):
            trainer.logged_metrics[f"lr-{type(opt).__name__}" + (f"-{i}" if i else "")] = opt.param_groups[0]["lr"]
