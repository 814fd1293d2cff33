"""`Trainer` stand-in: the optimisation loop Lightning 1.5 runs for the reference (main_id_embed.py:748,812), one
process per GPU.  Supported: fit() over train_dataloader with max_steps / max_epochs, accumulate_grad_batches, callbacks,
ModelCheckpoint cadence, save_checkpoint (same keys the reference reads back: "state_dict", "global_step", plus the
module's on_save_checkpoint additions), data-parallel gradient averaging over the trainable parameters through
celebbasis_b200.dist (torchrun environment).  Validation / test loops are no-ops (the reference runs with
--no-test and a dummy validation set)."""
import argparse
import os

import torch


class _Profiler:
    def summary(self):
        return ""


class _Plugin:
    def __init__(self, trainer):
        self.trainer = trainer

    def reduce(self, value, *args, **kwargs):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.as_tensor(float(value), device=self.trainer.device)
            dist.all_reduce(t)
            return (t / dist.get_world_size()).item()
        return value


_ARGS = {  # name -> (type, default): the flags the reference passes or compares (nondefault_trainer_args)
    "gpus": (str, None), "max_steps": (int, -1), "max_epochs": (int, None), "accumulate_grad_batches": (int, 1),
    "accelerator": (str, None), "resume_from_checkpoint": (str, None), "benchmark": (bool, False),
    "num_sanity_val_steps": (int, 2), "check_val_every_n_epoch": (int, 1), "val_check_interval": (float, 1.0),
    "log_every_n_steps": (int, 50), "precision": (int, 32), "num_nodes": (int, 1), "limit_val_batches": (float, 1.0),
    "default_root_dir": (str, None), "gradient_clip_val": (float, 0.0), "deterministic": (bool, False),
    "profiler": (str, None), "fast_dev_run": (bool, False), "find_unused_parameters": (bool, False),
}


def _move(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: _move(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_move(v, device) for v in obj)
    return obj


class Trainer:
    def __init__(self, gpus=None, max_steps=-1, max_epochs=None, accumulate_grad_batches=1, callbacks=None, logger=None,
                 default_root_dir=None, gradient_clip_val=0.0, resume_from_checkpoint=None, **kwargs):
        self.gpus = gpus
        self.max_steps = -1 if max_steps is None else int(max_steps)
        self.max_epochs = max_epochs
        self.accumulate_grad_batches = int(accumulate_grad_batches or 1)
        self.callbacks = list(callbacks or [])
        from ..callbacks import ModelCheckpoint
        self.checkpoint_callbacks = [c for c in self.callbacks if isinstance(c, ModelCheckpoint)]
        self.checkpoint_callback = self.checkpoint_callbacks[0] if self.checkpoint_callbacks else None
        self.logger = logger
        self.default_root_dir = default_root_dir or os.getcwd()
        self.gradient_clip_val = float(gradient_clip_val or 0.0)
        self.resume_from_checkpoint = resume_from_checkpoint
        if resume_from_checkpoint:
            import warnings
            warnings.warn("Trainer stand-in: resume_from_checkpoint is accepted for CLI compatibility but optimiser / "
                          "step state is not restored (the reference resumes weights through --actual_resume)")
        self.extra = kwargs
        self.global_step = 0
        self.current_epoch = 0
        self.optimizers = []
        self.lr_schedulers = []
        self.logged_metrics = {}
        self.interrupted = False
        self.profiler = _Profiler()
        self.training_type_plugin = _Plugin(self)
        self.lightning_module = None
        self.datamodule = None
        self.logdir = None
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        use_gpu = gpus not in (None, 0, "0", "") and torch.cuda.is_available()
        self.root_gpu = self.local_rank if use_gpu else None
        self.device = torch.device("cuda", self.local_rank) if use_gpu else torch.device("cpu")

    # ---- argparse plumbing (main_id_embed.py:183,540,748) -------------------------------------------------------
    @classmethod
    def add_argparse_args(cls, parser):
        for name, (typ, default) in _ARGS.items():
            if typ is bool:
                parser.add_argument(f"--{name}", type=lambda s: str(s).lower() in ("1", "true", "yes"), nargs="?",
                                    const=True, default=default)
            else:
                parser.add_argument(f"--{name}", type=typ, default=default)
        return parser

    @classmethod
    def from_argparse_args(cls, args, **kwargs):
        params = vars(args) if isinstance(args, argparse.Namespace) else dict(args)
        params = {k: v for k, v in params.items()}
        params.update(kwargs)
        return cls(**params)

    # ---- hooks ------------------------------------------------------------------------------------------------------
    def _module_hook(self, name, *args):
        fn = getattr(self.lightning_module, name, None)        # modules built on another LightningModule stand-in may lack it
        return fn(*args) if callable(fn) else None

    def _publish_progress(self):
        for name, val in (("global_step", self.global_step), ("current_epoch", self.current_epoch)):
            try:
                setattr(self.lightning_module, name, val)
            except AttributeError:                              # read-only property that already reads the trainer
                pass

    def _call(self, hook, *args):
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(self, self.lightning_module, *args)

    # ---- checkpointing ------------------------------------------------------------------------------------------------
    def save_checkpoint(self, filepath, weights_only=False):
        model = self.lightning_module
        # the ~4 GB CPU copy of the state dict is materialised only if somebody still wants it after the module's
        # on_save_checkpoint hook (the CelebBasis LatentDiffusion clears the checkpoint and writes embeddings.pt), rank 0 only
        ckpt = {"global_step": self.global_step, "epoch": self.current_epoch, "pytorch-lightning_version": "1.5.9",
                "state_dict": None}
        self._module_hook("on_save_checkpoint", ckpt)
        self._call("on_save_checkpoint", ckpt)
        if "state_dict" in ckpt and ckpt["state_dict"] is None:
            if self.global_rank == 0:
                ckpt["state_dict"] = {k: v.detach().cpu() for k, v in model.state_dict().items()}
                if not weights_only:
                    ckpt["optimizer_states"] = [o.state_dict() for o in self.optimizers]
            else:
                del ckpt["state_dict"]
        if self.global_rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(filepath)), exist_ok=True)
            torch.save(ckpt, filepath)
        return ckpt

    # ---- optimisation loop ----------------------------------------------------------------------------------------
    @staticmethod
    def _unpack_optimizers(conf):
        if isinstance(conf, dict):
            return [conf["optimizer"]], [conf["lr_scheduler"]] if "lr_scheduler" in conf else []
        if isinstance(conf, (list, tuple)):
            if len(conf) == 2 and isinstance(conf[0], (list, tuple)):
                return list(conf[0]), list(conf[1])
            return list(conf), []
        return [conf], []

    def _sync_grads(self, model):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        grads = [p.grad for p in model.parameters() if p.requires_grad and p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1).float() for g in grads])
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    def _lookahead(self, loader):
        """(index, batch on the device, next batch on the device or None): one batch of look-ahead, each batch moved once."""
        it = iter(loader)
        prep = getattr(self.lightning_module, "preprocess_batch", None)      # device-side data path, if the module has one
        prep = prep if callable(prep) else (lambda b: b)
        try:
            cur = prep(_move(next(it), self.device))
        except StopIteration:
            return
        idx = 0
        while True:
            try:
                nxt = prep(_move(next(it), self.device))
            except StopIteration:
                nxt = None
            yield idx, cur, nxt
            if nxt is None:
                return
            cur, idx = nxt, idx + 1

    def _shard_loader(self, loader):
        """Lightning's replace_sampler_ddp: with world_size > 1 every rank must see a disjoint shard of the dataset (the
        reference relies on it: main_id_embed.py seeds all ranks identically and scales the LR by ngpu, :778-779)."""
        if self.world_size <= 1:
            return loader
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler
        if not isinstance(loader, DataLoader) or isinstance(getattr(loader, "sampler", None), DistributedSampler):
            return loader
        if loader.batch_size is None:
            return loader
        from torch.utils.data import RandomSampler
        shuffle = isinstance(loader.sampler, RandomSampler)
        sampler = DistributedSampler(loader.dataset, num_replicas=self.world_size, rank=self.global_rank, shuffle=shuffle)
        return DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler, num_workers=loader.num_workers,
                          collate_fn=loader.collate_fn, pin_memory=loader.pin_memory, drop_last=loader.drop_last,
                          worker_init_fn=loader.worker_init_fn, persistent_workers=loader.persistent_workers
                          if loader.num_workers > 0 else False)

    def fit(self, model, datamodule=None, train_dataloaders=None):
        self.lightning_module = model
        model.trainer = self
        self.datamodule = datamodule
        if self.world_size > 1:
            from celebbasis_b200 import dist as cbd
            cbd.init()
        model.to(self.device)
        if datamodule is not None:
            datamodule.trainer = self
            if not hasattr(datamodule, "datasets"):
                datamodule.prepare_data()
                datamodule.setup("fit")
            loader = datamodule.train_dataloader()
        else:
            loader = train_dataloaders
        loader = self._shard_loader(loader)
        self.optimizers, self.lr_schedulers = self._unpack_optimizers(model.configure_optimizers())
        self._call("setup", "fit")
        self._call("on_pretrain_routine_start")
        model.train()
        self._call("on_train_start")
        done = False
        try:
            while not done:
                self._call("on_train_epoch_start")
                sampler = getattr(loader, "sampler", None)
                if hasattr(sampler, "set_epoch"):
                    sampler.set_epoch(self.current_epoch)      # Lightning does this for the DistributedSampler it installs
                for batch_idx, batch, nxt in self._lookahead(loader):
                    stage = getattr(model, "stage_next_batch", None)
                    if callable(stage):
                        stage(nxt)           # the module may overlap the next batch's frozen front end with this step
                    self._call("on_train_batch_start", batch, batch_idx, 0)
                    self._module_hook("on_train_batch_start", batch, batch_idx, 0)
                    out = model.training_step(batch, batch_idx)
                    loss = out["loss"] if isinstance(out, dict) else out
                    (loss / self.accumulate_grad_batches).backward()
                    if (batch_idx + 1) % self.accumulate_grad_batches == 0:
                        self._sync_grads(model)
                        if self.gradient_clip_val > 0:
                            torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None],
                                                           self.gradient_clip_val)
                        for opt in self.optimizers:
                            opt.step()
                            opt.zero_grad(set_to_none=True)
                        for sch in self.lr_schedulers:
                            (sch["scheduler"] if isinstance(sch, dict) else sch).step()
                        self.global_step += 1
                        self._publish_progress()
                    outputs = {"loss": loss.detach()}
                    self._module_hook("on_train_batch_end", outputs, batch, batch_idx, 0)
                    self._call("on_train_batch_end", outputs, batch, batch_idx, 0)
                    if self.max_steps is not None and 0 <= self.max_steps <= self.global_step:
                        done = True
                        break
                self._call("on_train_epoch_end")
                self.current_epoch += 1
                self._publish_progress()
                if self.max_epochs is not None and self.current_epoch >= self.max_epochs:
                    done = True
                if self.max_steps in (None, -1) and self.max_epochs is None and self.current_epoch >= 1000:
                    done = True      # Lightning's default max_epochs
        except KeyboardInterrupt:
            self.interrupted = True
            self._call("on_keyboard_interrupt")
        self._call("on_train_end")

    def validate(self, *args, **kwargs):
        return []

    def test(self, *args, **kwargs):
        return []
