"""Compatibility layer for running the reference's own driver scripts against this package (SURVEY.md 8b "third-party
surfaces that must exist for unchanged scripts", 8f-1): stand-ins for `omegaconf` and `pytorch_lightning==1.5.9`, and
import-only stubs for `taming`, `kornia`, `clip`, `test_tube`.  Nothing here is on the hot path.

    python -m celebbasis_b200.compat.run scripts/stable_txt2img.py --config ... --ckpt ...

registers the stand-ins for whatever is not installed, puts this repository's `ldm` mirror first on sys.path and runs
the script unchanged."""
import importlib
import importlib.util
import sys
import types

from torch import nn


def _missing(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _alias(real_name, shim_pkg):
    """Register shim package `shim_pkg` (and its already-imported submodules) under the real distribution's name."""
    mod = importlib.import_module(shim_pkg)
    sys.modules[real_name] = mod
    for sub in ("listconfig", "dictconfig", "core", "trainer", "callbacks", "loggers", "utilities", "utilities.distributed"):
        try:
            sys.modules[f"{real_name}.{sub}"] = importlib.import_module(f"{shim_pkg}.{sub}")
        except ImportError:
            pass
    return mod


def install(verbose=False):
    """Idempotent.  Returns the list of names that were provided by stand-ins."""
    provided = []
    if _missing("omegaconf"):
        _alias("omegaconf", "celebbasis_b200.compat.omegaconf")
        provided.append("omegaconf")
    if _missing("pytorch_lightning"):
        _alias("pytorch_lightning", "celebbasis_b200.compat.pytorch_lightning")
        provided.append("pytorch_lightning")
    if _missing("taming"):      # autoencoder.py:6 imports VectorQuantizer2 (never executed on this path)
        for name in ("taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2 = type("VectorQuantizer2", (nn.Module,), {})
        sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer = sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2
        provided.append("taming")
    for name in ("kornia", "clip", "test_tube"):            # modules.py:4,7 / meta_net.py:4: imported, not executed here
        if _missing(name):
            sys.modules[name] = types.ModuleType(name)
            provided.append(name)
    if verbose and provided:
        print(f"[celebbasis_b200.compat] stand-ins registered for: {', '.join(provided)}")
    return provided
