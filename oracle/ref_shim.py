"""ORACLE tooling (test infrastructure): import the UNMODIFIED reference from /root/reference and run it on CPU.

Used only by oracle/make_golden.py (in this container, where /root/reference exists) to pin
oracle/torch_ref.py and to generate tests/golden/*.  Nothing at run time on the GPU box imports this.

What is shimmed -- third-party packages the reference imports but that are not
installed here, and two constructors that need the network / a weights file:
  * omegaconf.listconfig.ListConfig, pytorch_lightning (LightningModule = nn.Module + device/log),
    taming...VectorQuantizer2, kornia, clip  ->  empty stand-ins (never executed on this path);
  * CLIPTokenizer.from_pretrained / CLIPTextModel.from_pretrained (modules.py:171-172) -> the deterministic
    SyntheticCLIPTokenizer below + a random-init CLIPTextModel of the requested depth;
  * the patched encoder_forward calls CLIPEncoderLayer with the transformers-4.18 signature
    (modules.py:325-330); transformers 5.5 layers take (hidden, mask): adapted by a loop with identical maths;
  * MetaIdNet.load_fr_net (meta_net.py:348-355) hard-codes a weights path -> iresnet100() with synthetic weights.
The arithmetic of every reference module on the path runs unmodified.
"""
import contextlib
import hashlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = os.environ.get("CELEBBASIS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from celebbasis_b200.tokenizer import BOS, EOS, KNOWN_TOKENS, SyntheticCLIPTokenizer  # noqa: E402


PHRASES = {}        # {text: real CLIP BPE ids}: set by make_golden.py for the celeb-basis fixture (infer_images/token_len.txt)


class _Cfg(dict):
    """dict with attribute access, enough of OmegaConf for ldm.util.instantiate_from_config / ddpm.py."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _Cfg(v) if isinstance(v, dict) and not isinstance(v, _Cfg) else v


def _wrap(d):
    if isinstance(d, dict):
        return _Cfg({k: _wrap(v) for k, v in d.items()})
    return d


def install_stubs(clip_layers=12):
    if "ldm" in sys.modules:
        paths = [str(p) for p in getattr(sys.modules["ldm"], "__path__", [])]
        if not any(p.startswith(REF_ROOT) for p in paths):
            raise RuntimeError("a different `ldm` package is already imported; run the oracle in its own process")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    if "ldm" not in sys.modules:
        # The reference's `ldm` directory has no __init__.py (namespace package), so any regular package named `ldm` on
        # sys.path -- this repository's drop-in mirror -- would win regardless of order: bind the name to the reference
        # tree explicitly.
        import importlib.machinery
        import importlib.util
        spec = importlib.machinery.ModuleSpec("ldm", None, is_package=True)
        spec.submodule_search_locations = [os.path.join(REF_ROOT, "ldm")]
        sys.modules["ldm"] = importlib.util.module_from_spec(spec)

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    oc = mod("omegaconf")
    lc = mod("omegaconf.listconfig")

    class ListConfig(list):
        pass
    lc.ListConfig = ListConfig
    oc.listconfig = lc
    oc.OmegaConf = type("OmegaConf", (), {})

    pl = mod("pytorch_lightning")

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.global_step = 0
            self.current_epoch = 0

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass
    pl.LightningModule = LightningModule
    plu = mod("pytorch_lightning.utilities")
    plud = mod("pytorch_lightning.utilities.distributed")
    plud.rank_zero_only = lambda f: f
    plu.distributed = plud
    pl.utilities = plu

    for name in ["taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"]:
        mod(name)
    sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2 = type("VectorQuantizer2", (nn.Module,), {})
    mod("kornia")
    mod("clip")

    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTokenizer

    def tok_from_pretrained(*a, **k):
        return SyntheticCLIPTokenizer(phrases=PHRASES)

    def model_from_pretrained(*a, **k):
        cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=clip_layers,
                             num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                             layer_norm_eps=1e-5, attn_implementation="eager")
        return CLIPTextModel(cfg)

    CLIPTokenizer.from_pretrained = staticmethod(tok_from_pretrained)
    CLIPTextModel.from_pretrained = staticmethod(model_from_pretrained)


def _adapt_clip_encoder(embedder):
    """Replace the 4.18-signature encoder loop (modules.py:302-342) by the same loop for transformers>=5."""
    enc = embedder.transformer.text_model.encoder

    def encoder_forward(inputs_embeds, attention_mask=None, causal_attention_mask=None, **_):
        mask = causal_attention_mask if attention_mask is None else causal_attention_mask + attention_mask
        h = inputs_embeds
        for layer in enc.layers:
            h = layer(h, mask)
        return h
    enc.forward = encoder_forward


def build_reference(model_params, seed=0, clip_layers=12, celeb_basis=None):
    """Construct the reference LatentDiffusion (CPU) from a params dict shaped like configs/.../aigc_id.yaml:model.params,
    load the synthetic checkpoint, and return it in the state main_id_embed.py would train it in."""
    install_stubs(clip_layers)
    from celebbasis_b200 import synth
    import ldm.modules.id_embedding.meta_net as meta_net
    from ldm.modules.id_embedding.iresnet import iresnet100

    def load_fr_net(self):
        self.id_model = iresnet100()
        for p in self.id_model.parameters():
            p.requires_grad = False
        self.id_model.eval()
    meta_net.MetaIdNet.load_fr_net = load_fr_net

    from ldm.models.diffusion.ddpm import LatentDiffusion
    params = _wrap(model_params)
    model = LatentDiffusion(**params)
    _adapt_clip_encoder(model.cond_stage_model)
    sd = synth.synth_state_dict(model, seed=seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    if celeb_basis is not None:
        model.cond_stage_model.celeb_embeddings = celeb_basis.clone()
    model.train()  # EmbeddingManagerId switches on self.training; UNet/CLIP/VAE stay eval via disabled_train
    return model


@contextlib.contextmanager
def replay_randomness(t, noise, posterior_eps):
    """Feed ddpm.py:927 (randint), ddpm.py:1070 (randn_like) and distributions.py:36 (randn) from given tensors."""
    o_randint, o_randn_like, o_randn = torch.randint, torch.randn_like, torch.randn
    state = {"eps_used": False}

    def randint(*a, **k):
        return t.clone()

    def randn_like(x, *a, **k):
        assert x.shape == noise.shape, (x.shape, noise.shape)
        return noise.clone().to(x.dtype)

    def randn(*shape, **k):
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        if shp == tuple(posterior_eps.shape) and not state["eps_used"]:
            state["eps_used"] = True
            return posterior_eps.clone()
        return o_randn(*shape, **k)
    torch.randint, torch.randn_like, torch.randn = randint, randn_like, randn
    try:
        yield
    finally:
        torch.randint, torch.randn_like, torch.randn = o_randint, o_randn_like, o_randn
