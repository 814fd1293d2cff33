"""Mirror of the helpers of ldm/util.py the hot path uses (instantiate_from_config :78-93, default/exists :55-64)."""
import importlib
from inspect import isfunction


def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if isfunction(d) else d


def count_params(model, verbose=False):
    total = sum(p.numel() for p in model.parameters())
    if verbose:
        print(f"{model.__class__.__name__} has {total * 1.e-6:.2f} M params.")
    return total


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config, **kwargs):
    if "target" not in config:
        if config == '__is_first_stage__' or config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()), **kwargs)


def cfg_get(cfg, *keys, default=None):
    """Nested lookup that works for plain dicts and OmegaConf nodes alike."""
    cur = cfg
    for k in keys:
        try:
            cur = cur[k]
        except Exception:
            return default
    return cur


def log_txt_as_img(wh, xc, size=10):
    """ldm/util.py:17-38: the captions rendered as (B,3,H,W) images in [-1,1] for the image logger (host, PIL)."""
    import numpy as np
    import torch
    from PIL import Image, ImageDraw, ImageFont
    txts = []
    for cap in xc:
        txt = Image.new("RGB", wh, color="white")
        draw = ImageDraw.Draw(txt)
        nc = max(1, int(40 * (wh[0] / 256)))
        lines = "\n".join(cap[start:start + nc] for start in range(0, len(cap), nc))
        try:
            draw.text((0, 0), lines, fill="black", font=ImageFont.load_default())
        except UnicodeEncodeError:
            print("Cant encode string for logging. Skipping.")
        txts.append(np.array(txt).transpose(2, 0, 1) / 127.5 - 1.0)
    return torch.tensor(np.stack(txts))
