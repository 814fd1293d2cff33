"""Drop-in import path: `ldm.*` resolves to the B200-native host mirror in celebbasis_b200/ldm.

The reference's configs name their classes by import path (configs/stable-diffusion/aigc_id.yaml:3,24,40,57,80:
ldm.models.diffusion.ddpm.LatentDiffusion, ldm.modules.diffusionmodules.openaimodel.UNetModel, ...), and
main.py / scripts/stable_txt2img.py import them the same way, so the mirror must be importable as `ldm`.
"""
import os as _os

import celebbasis_b200.ldm as _impl

__path__ = [_os.path.dirname(_impl.__file__)]
