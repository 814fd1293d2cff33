"""Host mirror of ldm/models/diffusion/ddpm.py (DDPM :46, LatentDiffusion :439, DiffusionWrapper :1530).

Keeps the surface main.py / main_id_embed.py / scripts/stable_txt2img.py / the DDIM sampler use:
constructor keywords of configs/stable-diffusion/aigc_id.yaml, the schedule buffers, `shared_step`, `forward`,
`p_losses`, `apply_model`, `get_input`, `get_learned_conditioning`, `encode/decode_first_stage`, `q_sample`,
`configure_optimizers`, `training_step`, `on_save_checkpoint`, `ema_scope`.  Every tensor operation of the step is a
kernel of libcelebbasis_b200.so reached through the mirrored sub-modules; this file is glue, exactly as in the
reference.  Lightning is optional: without pytorch_lightning the class derives from a minimal stand-in.
"""
import os
from contextlib import contextmanager
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from celebbasis_b200 import ops
from ldm.modules.diffusionmodules.util import extract_into_tensor, make_beta_schedule, noise_like
from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
from ldm.util import cfg_get, count_params, default, exists, instantiate_from_config

try:  # pragma: no cover - pytorch_lightning is not installed in this image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    from pytorch_lightning.utilities.distributed import rank_zero_only
except Exception:  # minimal LightningModule protocol (what ddpm.py touches)
    class _Base(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.global_step = 0
            self.current_epoch = 0
            self.trainer = None

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, d, *a, **k):
            self.last_log = dict(d)

    def rank_zero_only(fn):
        return fn

__conditioning_keys__ = {'concat': 'c_concat', 'crossattn': 'c_crossattn', 'adm': 'y'}


def disabled_train(self, mode=True):
    return self


class _MSEFn(torch.autograd.Function):
    """loss_simple[b] = mean_(c,h,w) (pred - target)^2  (get_loss 'l2' + .mean([1,2,3]), ddpm.py:294-307,1084)."""

    @staticmethod
    def forward(ctx, pred, target):
        need = ctx.needs_input_grad[0]     # (grad mode is off inside Function.forward; this reflects the call site)
        loss, grad = ops.mse_fwd_bwd(pred.float().contiguous(), target.float().contiguous(), 1.0, want_grad=need)
        ctx.B = pred.shape[0]
        if need:
            ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors                 # = d(mean_b loss_b)/dpred ; chain rule with dloss per sample
        out = torch.empty_like(grad)
        w = (dloss.float() * ctx.B).tolist() if ctx.B > 1 else None
        if w is None:
            ops.axpby(grad.view(1, -1), float(dloss.item()), out=out.view(1, -1))
        else:
            for b in range(ctx.B):
                ops.axpby(grad[b].view(1, -1), w[b], out=out[b].view(1, -1))
        return out, None


def _plain(cfg):
    """OmegaConf node / dict / list -> plain python containers."""
    if isinstance(cfg, dict):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg


class _FusedLossFn(torch.autograd.Function):
    """The fused step computes loss AND d(loss)/d(W, b) in one CUDA-graph replay; this node hands those gradients to
    autograd so that the reference's `loss.backward(); optimizer.step()` contract (Lightning's loop) is unchanged."""

    @staticmethod
    def forward(ctx, W, b, loss_dev, gW, gb):
        ctx.save_for_backward(gW, gb)
        return loss_dev.reshape(()).clone()

    @staticmethod
    def backward(ctx, dloss):
        gW, gb = ctx.saved_tensors
        return gW * dloss, gb * dloss, None, None, None


class DDPM(_Base):
    def __init__(self, unet_config, timesteps=1000, beta_schedule="linear", loss_type="l2", ckpt_path=None,
                 ignore_keys=[], load_only_unet=False, monitor="val/loss", use_ema=True, first_stage_key="image",
                 image_size=256, channels=3, log_every_t=100, clip_denoised=True, linear_start=1e-4, linear_end=2e-2,
                 cosine_s=8e-3, given_betas=None, original_elbo_weight=0., embedding_reg_weight=0.,
                 unfreeze_model=False, model_lr=0., v_posterior=0., l_simple_weight=1., conditioning_key=None,
                 parameterization="eps", scheduler_config=None, use_positional_encodings=False, learn_logvar=False,
                 logvar_init=0.):
        super().__init__()
        assert parameterization == "eps", "SD-v1 / CelebBasis is eps-prediction"
        self.parameterization = parameterization
        self.cond_stage_model = None
        self.clip_denoised = clip_denoised
        self.log_every_t = log_every_t
        self.first_stage_key = first_stage_key
        self.image_size = image_size
        self.channels = channels
        self.use_positional_encodings = use_positional_encodings
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        count_params(self.model, verbose=True)
        self.use_ema = use_ema
        assert not use_ema, "use_ema: False in every CelebBasis config (aigc_id.yaml:18)"
        self.use_scheduler = scheduler_config is not None
        self.v_posterior = v_posterior
        self.original_elbo_weight = original_elbo_weight
        self.l_simple_weight = l_simple_weight
        self.embedding_reg_weight = embedding_reg_weight
        self.unfreeze_model = unfreeze_model
        self.model_lr = model_lr
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys, only_model=load_only_unet)
        self.register_schedule(given_betas=given_betas, beta_schedule=beta_schedule, timesteps=timesteps,
                               linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        self.loss_type = loss_type
        self.learn_logvar = learn_logvar
        self.logvar = torch.full(fill_value=logvar_init, size=(self.num_timesteps,))
        self.learning_rate = 5.0e-03

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """ddpm.py:126-178: float64 numpy schedule -> fp32 buffers (host arithmetic, init time)."""
        betas = given_betas if exists(given_betas) else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        to_torch = partial(torch.tensor, dtype=torch.float32)
        self.register_buffer('betas', to_torch(betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod)))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - alphas_cumprod)))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod - 1)))
        posterior_variance = (1 - self.v_posterior) * betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod) \
            + self.v_posterior * betas
        self.register_buffer('posterior_variance', to_torch(posterior_variance))
        self.register_buffer('posterior_log_variance_clipped', to_torch(np.log(np.maximum(posterior_variance, 1e-20))))
        self.register_buffer('posterior_mean_coef1', to_torch(betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod)))
        self.register_buffer('posterior_mean_coef2', to_torch((1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod)))
        lvlb = self.betas ** 2 / (2 * self.posterior_variance * to_torch(alphas) * (1 - self.alphas_cumprod))
        lvlb[0] = lvlb[1]
        self.register_buffer('lvlb_weights', lvlb, persistent=False)

    @contextmanager
    def ema_scope(self, context=None):
        yield None   # use_ema is False: the reference's scope is a no-op too (ddpm.py:180-194)

    def init_from_ckpt(self, path, ignore_keys=list(), only_model=False):
        sd = torch.load(path, map_location="cpu")
        if "state_dict" in list(sd.keys()):
            sd = sd["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        missing, unexpected = self.load_state_dict(sd, strict=False) if not only_model else self.model.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    def q_sample(self, x_start, t, noise=None):
        """ddpm.py:289-292 on the device (cb_q_sample)."""
        noise = default(noise, lambda: torch.randn_like(x_start))
        return ops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), t.long().contiguous(),
                            self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)

    def get_loss(self, pred, target, mean=True):
        assert self.loss_type == 'l2'
        per_sample = _MSEFn.apply(pred, target)
        return per_sample.mean() if mean else per_sample

    def get_input(self, batch, k):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        # 'b h w c -> b c h w' + contiguous + float (ddpm.py:344-350): pure layout glue
        return x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format).float()

    def training_step(self, batch, batch_idx):
        loss, loss_dict = self.shared_step(batch)
        self.log_dict(loss_dict, prog_bar=True, logger=True, on_step=True, on_epoch=True)
        self.log("global_step", self.global_step, prog_bar=True, logger=True, on_step=True, on_epoch=False)
        return loss


class LatentDiffusion(DDPM):
    """main class"""

    def __init__(self, first_stage_config, cond_stage_config, personalization_config, num_timesteps_cond=None,
                 cond_stage_key="image", cond_stage_trainable=False, concat_mode=True, cond_stage_forward=None,
                 conditioning_key=None, scale_factor=1.0, scale_by_std=False, *args, **kwargs):
        self.num_timesteps_cond = default(num_timesteps_cond, 1)
        self.scale_by_std = scale_by_std
        self._engine_params = _plain(dict(
            unet_config=kwargs.get("unet_config"), first_stage_config=first_stage_config,
            personalization_config=personalization_config, scale_factor=scale_factor,
            timesteps=kwargs.get("timesteps", 1000), linear_start=kwargs.get("linear_start", 1e-4),
            linear_end=kwargs.get("linear_end", 2e-2)))
        self._fused = None            # StepGraphs of the CUDA-graph training step (built on first use)
        self._fused_sig = None
        self._staged_next = None      # look-ahead batch handed over by the training loop (stage_next_batch)
        self.fused_step = os.environ.get("CB_FUSED_STEP", "1") != "0"
        assert self.num_timesteps_cond <= kwargs['timesteps']
        if conditioning_key is None:
            conditioning_key = 'concat' if concat_mode else 'crossattn'
        ckpt_path = kwargs.pop("ckpt_path", None)
        ignore_keys = kwargs.pop("ignore_keys", [])
        super().__init__(conditioning_key=conditioning_key, *args, **kwargs)
        self.concat_mode = concat_mode
        self.cond_stage_trainable = cond_stage_trainable
        self.cond_stage_key = cond_stage_key
        try:
            self.num_downs = len(cfg_get(first_stage_config, "params", "ddconfig", "ch_mult")) - 1
        except Exception:
            self.num_downs = 0
        assert not scale_by_std
        self.scale_factor = scale_factor
        self.instantiate_first_stage(first_stage_config)
        self.instantiate_cond_stage(cond_stage_config)
        self.cond_stage_forward = cond_stage_forward
        self.clip_denoised = False
        self.bbox_tokenizer = None
        self.restarted_from_ckpt = False
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys)
            self.restarted_from_ckpt = True
        if not self.unfreeze_model:
            self.cond_stage_model.eval()
            self.cond_stage_model.train = disabled_train
            for param in self.cond_stage_model.parameters():
                param.requires_grad = False
            self.model.eval()
            self.model.train = disabled_train
            for param in self.model.parameters():
                param.requires_grad = False
        self.embedding_manager = self.instantiate_embedding_manager(personalization_config, self.cond_stage_model)
        for param in self.embedding_manager.embedding_parameters():
            param.requires_grad = True
        for param in self.embedding_manager.trainable_parameters():
            param.requires_grad = True

    # ---- sub-module construction (ddpm.py:540-576) -----------------------------------------------------------
    def instantiate_first_stage(self, config):
        model = instantiate_from_config(config)
        self.first_stage_model = model.eval()
        self.first_stage_model.train = disabled_train
        for param in self.first_stage_model.parameters():
            param.requires_grad = False

    def instantiate_cond_stage(self, config):
        model = instantiate_from_config(config)
        if not self.cond_stage_trainable:
            self.cond_stage_model = model.eval()
            self.cond_stage_model.train = disabled_train
            for param in self.cond_stage_model.parameters():
                param.requires_grad = False
        else:
            self.cond_stage_model = model

    def instantiate_embedding_manager(self, config, embedder):
        model = instantiate_from_config(config, embedder=embedder)
        ckpt = cfg_get(config, "params", "embedding_manager_ckpt")
        if ckpt:
            model.load(ckpt)
        return model

    # ---- first stage ---------------------------------------------------------------------------------------------
    def get_first_stage_encoding(self, encoder_posterior):
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            return encoder_posterior.sample(scale=self.scale_factor)     # scale fused into the sampling kernel
        elif isinstance(encoder_posterior, torch.Tensor):
            return ops.axpby(encoder_posterior.reshape(encoder_posterior.shape[0], -1).float().contiguous(),
                             float(self.scale_factor)).view(encoder_posterior.shape)
        raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")

    @torch.no_grad()
    def encode_first_stage(self, x):
        return self.first_stage_model.encode(x)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        z = ops.axpby(z.reshape(z.shape[0], -1).float().contiguous(), 1. / self.scale_factor).view(z.shape)
        return self.first_stage_model.decode(z)

    # ---- conditioning ---------------------------------------------------------------------------------------------
    def get_learned_conditioning(self, c, face_img=None, image_ori=None):
        assert self.cond_stage_forward is None
        c = self.cond_stage_model.encode(c, embedding_manager=self.embedding_manager, face_img=face_img,
                                         image_ori=image_ori)
        if isinstance(c, DiagonalGaussianDistribution):
            c = c.mode()
        return c

    @torch.no_grad()
    def get_input(self, batch, k, return_first_stage_outputs=False, force_c_encode=False, cond_key=None,
                  return_original_cond=False, bs=None):
        x = super().get_input(batch, k)
        if bs is not None:
            x = x[:bs]
        x = x.to(self.device)
        encoder_posterior = self.encode_first_stage(x)
        z = self.get_first_stage_encoding(encoder_posterior).detach()
        cond_key = cond_key or self.cond_stage_key
        assert cond_key in ['caption', 'coordinates_bbox']
        xc = batch[cond_key]
        if not self.cond_stage_trainable or force_c_encode:
            c = self.get_learned_conditioning(xc, face_img=batch.get('image'), image_ori=batch.get('image_ori'))
        else:
            c = xc
        if bs is not None:
            c = c[:bs]
        c = {'caption': c, 'image': batch['image'], 'image_ori': batch.get('image_ori')}
        out = [z, c]
        if return_first_stage_outputs:
            out.extend([x, self.decode_first_stage(z)])
        if return_original_cond:
            out.append(xc)
        return out

    # ---- the training step (ddpm.py:921-936,948-1049,1069-1116) ---------------------------------------------------
    def preprocess_batch(self, batch):
        """RAW batches of the ldm.data.face_id mirror (uint8 images + drawn augmentation parameters) -> the batch dict the
        reference's DataLoader yields, with the pixel work done on this module's device (celebbasis_b200.data_path)."""
        from celebbasis_b200 import data_path
        if data_path.is_raw_batch(batch):
            return data_path.device_augment(batch, self.device)
        return batch

    def shared_step(self, batch, **kwargs):
        batch = self.preprocess_batch(batch)
        if self._fused_applicable(batch):
            return self._fused_shared_step(batch)
        x, c = self.get_input(batch, self.first_stage_key)
        return self(x, c['caption'], face_img=c['image'], image_ori=c['image_ori'])

    # ---- fused training step: the whole of shared_step + backward as CUDA-graph replays ---------------------------
    def stage_next_batch(self, batch):
        """Optional look-ahead hook for the training loop: the batch that the NEXT training_step call will receive (the
        same object).  Its frozen front end (VAE encode, CosFace features) then overlaps this step's UNet work."""
        self._staged_next = batch

    def _fused_applicable(self, batch):
        if not (self.fused_step and self.training and self.cond_stage_trainable and not self.unfreeze_model):
            return False
        if self.embedding_reg_weight > 0 or self.original_elbo_weight != 0 or self.l_simple_weight != 1.:
            return False
        io = batch.get("image_ori") if isinstance(batch, dict) else None
        x = batch.get(self.first_stage_key) if isinstance(batch, dict) else None
        if io is None or not torch.is_tensor(x) or x.dim() != 4 or x.shape[-1] != 3 or x.dtype != torch.float32:
            return False
        faces, nid = io.get("faces"), io.get("num_ids")
        if not torch.is_tensor(faces) or faces.shape[:3] != x.shape[:3] or faces.shape[-1] % 3 != 0:
            return False
        if nid is None or not bool((torch.as_tensor(nid).cpu() == 1).all()):
            return False           # two / three persons per prompt: the eager per-module path handles them
        if getattr(self.cond_stage_model, "celeb_embeddings", None) is None:
            return False
        return next(self.model.parameters()).is_cuda

    def _fused_build(self, batch):
        from celebbasis_b200.step_graph import StepGraphs
        from celebbasis_b200.train_step import CelebBasisStep
        x = batch[self.first_stage_key]
        faces = batch["image_ori"]["faces"]
        B, hw, n_chunks = x.shape[0], x.shape[1], faces.shape[-1] // 3
        em = self.embedding_manager
        lin = em.meta_id_net.stylegan_mlp.net[0]
        dev = lin.weight.device
        eng = CelebBasisStep(self._engine_params, self.state_dict(), self.cond_stage_model.celeb_embeddings, dev,
                             tokenizer=self.cond_stage_model.tokenizer, placeholder=em.placeholder_strings[0],
                             lr=self.learning_rate, id_coefficients=em.id_coefficients, id_embeddings=em.id_embeddings)
        # one storage for the trainable tensors and the per-identity EMA state: the optimiser (FusedAdamW on the mirror's
        # parameters), save()/load() of the embedding manager and the graph all see the same memory
        eng.flat[: lin.weight.numel()].copy_(lin.weight.detach().reshape(-1))
        eng.flat[lin.weight.numel():].copy_(lin.bias.detach().reshape(-1))
        lin.weight.data, lin.bias.data = eng.W, eng.b
        em.id_coefficients = list(eng.id_coefficients.unbind(0))
        em.id_embeddings = list(eng.id_embeddings.unbind(0))
        em.moved_to_device = True
        T = getattr(self.cond_stage_model, "max_length", 77)
        G = StepGraphs(eng, B=B, T=T, n_chunks=n_chunks, image_hw=hw)
        ids, map_np, _ = eng.prepare(batch["caption"])
        lat = G.noise.shape[-1]
        G.load_next(x, faces, torch.zeros(B, 4, lat, lat))
        G.load_step(ids, map_np, torch.zeros(B, dtype=torch.long), torch.zeros_like(G.noise), batch["image_ori"]["ids"])
        G.capture()
        self._fused, self._fused_sig = G, (B, hw, n_chunks, dev)
        return G

    def _fused_shared_step(self, batch):
        x = batch[self.first_stage_key]
        io = batch["image_ori"]
        B, hw = x.shape[0], x.shape[1]
        G = self._fused
        dev = next(self.model.parameters()).device
        if G is None or self._fused_sig != (B, hw, io["faces"].shape[-1] // 3, dev):
            G = self._fused_build(batch)
        eng = G.eng
        lat_shape = list(G.peps_n.shape)
        if G.next_token is not batch:                         # no look-ahead happened for this batch: front end now
            G.load_next(x, io["faces"], torch.randn(lat_shape))       # posterior eps from the CPU generator, as
            G.prefetch(batch)                                          # distributions.py:36 does
        ids, map_np, positions = eng.prepare(batch["caption"])        # tokenise + bit-exact placeholder row map (host)
        self.embedding_manager.last_positions = positions
        t = torch.randint(0, self.num_timesteps, (B,), device=dev).long()
        noise = torch.randn_like(G.z)
        G.load_step(ids, map_np, t, noise, io["ids"])
        nxt, self._staged_next = self._staged_next, None
        if nxt is not None and (nxt is batch or not self._fused_applicable(nxt)
                                or nxt[self.first_stage_key].shape != x.shape):
            nxt = None
        if nxt is not None:
            G.load_next(nxt[self.first_stage_key], nxt["image_ori"]["faces"], torch.randn(lat_shape))
        loss_dev = G.step(lookahead=nxt is not None, token=nxt)
        lin = self.embedding_manager.meta_id_net.stylegan_mlp.net[0]
        loss = _FusedLossFn.apply(lin.weight, lin.bias, loss_dev, eng.gW, eng.gb)
        loss_simple = eng.last["loss_simple"].detach()
        prefix = 'train' if self.training else 'val'
        loss_dict = {f'{prefix}/loss_simple': loss_simple.mean(),
                     f'{prefix}/loss_vlb': (self.lvlb_weights.to(dev)[t] * loss_simple).mean(),
                     f'{prefix}/loss_emb_reg': self.embedding_manager.embedding_neg_loss(),
                     f'{prefix}/loss': loss.detach()}
        return loss, loss_dict

    def forward(self, x, c, face_img=None, image_ori=None, *args, **kwargs):
        t = torch.randint(0, self.num_timesteps, (x.shape[0],), device=self.device).long()
        if self.model.conditioning_key is not None:
            assert c is not None
            if self.cond_stage_trainable:
                c = self.get_learned_conditioning(c, face_img=face_img, image_ori=image_ori)
        return self.p_losses(x, c, t, *args, **kwargs)

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = 'c_concat' if self.model.conditioning_key == 'concat' else 'c_crossattn'
            cond = {key: cond}
        x_recon = self.model(x_noisy, t, **cond)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    def p_losses(self, x_start, cond, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        model_output = self.apply_model(x_noisy, t, cond)
        loss_dict = {}
        prefix = 'train' if self.training else 'val'
        target = noise
        loss_simple = self.get_loss(model_output, target, mean=False)          # (B,) already averaged over (C,H,W)
        loss_dict.update({f'{prefix}/loss_simple': loss_simple.mean()})
        if self.logvar.device != t.device:
            self.logvar = self.logvar.to(t.device)
        logvar_t = self.logvar[t]
        loss = loss_simple / torch.exp(logvar_t) + logvar_t
        loss = self.l_simple_weight * loss.mean()
        loss_vlb = (self.lvlb_weights[t] * loss_simple.detach()).mean()
        loss_dict.update({f'{prefix}/loss_vlb': loss_vlb})
        loss = loss + self.original_elbo_weight * loss_vlb
        loss_dict.update({f'{prefix}/loss': loss})
        if self.embedding_reg_weight > 0:
            reg = self.embedding_manager.embedding_to_coarse_loss()
            loss_dict.update({f'{prefix}/loss_emb_reg': reg})
            loss = loss + self.embedding_reg_weight * reg
        neg = self.embedding_manager.embedding_neg_loss()
        loss = loss + neg * 1.
        loss_dict.update({f'{prefix}/loss_emb_reg': neg})
        loss_dict.update({f'{prefix}/loss': loss})
        return loss, loss_dict

    # ---- image logging (ddpm.py:1305-1440; called by main.ImageLogger every batch_frequency steps) --------------------
    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, **kwargs):
        assert ddim, "the CelebBasis configs log with the DDIM sampler"
        from ldm.models.diffusion.ddim import DDIMSampler
        shape = (self.channels, self.image_size, self.image_size)
        return DDIMSampler(self).sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)

    @torch.no_grad()
    def log_images(self, batch, N=8, n_row=4, sample=True, ddim_steps=50, ddim_eta=1., return_keys=None,
                   quantize_denoised=True, inpaint=False, plot_denoise_rows=False, plot_progressive_rows=False,
                   plot_diffusion_rows=False, **kwargs):
        """inputs / reconstruction / rendered captions / DDIM samples (plain and with guidance 5.0), as the reference logs
        them (the inpainting, progressive and diffusion-row panels are off in every CelebBasis config)."""
        from ldm.util import log_txt_as_img
        assert not (inpaint or plot_denoise_rows or plot_progressive_rows or plot_diffusion_rows)
        batch = self.preprocess_batch(batch)
        log = dict()
        z, c, x, xrec, xc = self.get_input(batch, self.first_stage_key, return_first_stage_outputs=True,
                                           force_c_encode=True, return_original_cond=True, bs=N)
        c = c['caption']
        N = min(x.shape[0], N)
        log["inputs"] = x
        log["reconstruction"] = xrec
        if self.model.conditioning_key is not None and self.cond_stage_key in ["caption"]:
            log["conditioning"] = log_txt_as_img((x.shape[2], x.shape[3]), batch["caption"][:N])
        if sample:
            with self.ema_scope("Plotting"):
                samples, _ = self.sample_log(cond=c, batch_size=N, ddim=ddim_steps is not None, ddim_steps=ddim_steps,
                                             eta=ddim_eta)
            log["samples"] = self.decode_first_stage(samples)
            uc = self.get_learned_conditioning(len(c) * [""])
            sample_scaled, _ = self.sample_log(cond=c, batch_size=N, ddim=ddim_steps is not None, ddim_steps=ddim_steps,
                                               eta=ddim_eta, unconditional_guidance_scale=5.0,
                                               unconditional_conditioning=uc)
            log["samples_scaled"] = self.decode_first_stage(sample_scaled)
        if return_keys:
            if np.intersect1d(list(log.keys()), return_keys).shape[0] == 0:
                return log
            return {key: log[key] for key in return_keys}
        return log

    # ---- optimisation / checkpoint cadence (ddpm.py:1442-1454,1519-1528) --------------------------------------------
    def configure_optimizers(self):
        lr = self.learning_rate
        params = list(self.embedding_manager.embedding_parameters()) + list(self.embedding_manager.trainable_parameters())
        from celebbasis_b200.optim import FusedAdamW
        return FusedAdamW([p for p in params if p.requires_grad], lr=lr)

    @rank_zero_only
    def on_save_checkpoint(self, checkpoint):
        checkpoint.clear()
        logdir = getattr(getattr(self, "trainer", None), "checkpoint_callback", None)
        dirpath = getattr(logdir, "dirpath", None) or "."
        if os.path.isdir(dirpath):
            self.embedding_manager.save(os.path.join(dirpath, "embeddings.pt"))
            self.embedding_manager.save(os.path.join(dirpath, f"embeddings_gs-{self.global_step}.pt"))


class DiffusionWrapper(_Base):
    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, 'crossattn']

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None):
        if self.conditioning_key is None:
            return self.diffusion_model(x, t)
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc)
