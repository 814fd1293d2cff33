"""Minimal in-repo stand-in for the `omegaconf` surface the reference drivers touch (`OmegaConf.load /
create / merge / from_dotlist / to_yaml / to_container / save`, attribute + item access, `.get`, `.pop`, `in`,
`omegaconf.listconfig.ListConfig`): main_id_embed.py:590-595,640-739, scripts/stable_txt2img.py:228,
ldm/util.py:78-93, openaimodel.py:476.  Used only when the real package is not installed (compat.install())."""
import copy

import yaml

from .listconfig import ListConfig


class DictConfig(dict):
    """dict with attribute access; nested dicts / lists are wrapped on the way in."""

    def __init__(self, content=None):
        super().__init__()
        for k, v in (content or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, _wrap(v))

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"Missing key {k}") from None

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def update(self, other=(), **kw):
        for k, v in dict(other, **kw).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def copy(self):
        return DictConfig(self)

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, (DictConfig, ListConfig)):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig([_wrap(x) for x in v])
    return v


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


def _merge_into(dst, src):
    for k, v in src.items():
        if k in dst and isinstance(dst[k], dict) and isinstance(v, dict):
            _merge_into(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)


def _parse_scalar(text):
    try:
        v = yaml.safe_load(text)
    except yaml.YAMLError:
        return text
    if isinstance(v, str):              # YAML 1.1 reads "1e-3" as a string; OmegaConf reads a float
        try:
            return float(v)
        except ValueError:
            return v
    return v


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if obj is None:
            return DictConfig()
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return _wrap(copy.deepcopy(_plain(obj)))

    @staticmethod
    def load(path_or_file):
        if hasattr(path_or_file, "read"):
            return OmegaConf.create(yaml.safe_load(path_or_file) or {})
        with open(path_or_file) as f:
            return OmegaConf.create(yaml.safe_load(f) or {})

    @staticmethod
    def merge(*configs):
        out = DictConfig()
        for c in configs:
            if c is None:
                continue
            _merge_into(out, c if isinstance(c, dict) else OmegaConf.create(c))
        return out

    @staticmethod
    def from_dotlist(dotlist):
        out = DictConfig()
        for item in dotlist:
            key, _, val = item.partition("=")
            node = out
            parts = key.lstrip("-").split(".")
            for p in parts[:-1]:
                node = node.setdefault(p, DictConfig())
            node[parts[-1]] = _parse_scalar(val) if val != "" else None
        return out

    @staticmethod
    def to_container(cfg, resolve=True):
        return _plain(cfg)

    @staticmethod
    def to_yaml(cfg):
        return yaml.safe_dump(_plain(cfg), default_flow_style=False, sort_keys=False)

    @staticmethod
    def save(config, f):
        text = OmegaConf.to_yaml(config)
        if hasattr(f, "write"):
            f.write(text)
        else:
            with open(f, "w") as fh:
                fh.write(text)

    @staticmethod
    def select(cfg, key, default=None):
        node = cfg
        for p in key.split("."):
            if not isinstance(node, dict) or p not in node:
                return default
            node = node[p]
        return node

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (DictConfig, ListConfig))


__all__ = ["OmegaConf", "DictConfig", "ListConfig"]
