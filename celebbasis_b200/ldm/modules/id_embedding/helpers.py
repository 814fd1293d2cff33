"""Host mirror of ldm/modules/id_embedding/helpers.py:6-41 (integer path, bit-exact)."""
from typing import List

import numpy as np
import torch

from celebbasis_b200.train_step import get_rep_pos as _get_rep_pos_np, placeholder_row_map


def get_rep_pos(tokenized: torch.Tensor, rep_tokens: list):
    tok = tokenized.detach().cpu().numpy() if isinstance(tokenized, torch.Tensor) else np.asarray(tokenized)
    return _get_rep_pos_np(tok, [int(t) for t in rep_tokens])


def shift_tensor_dim0(ori: torch.Tensor, r_pos: List[np.ndarray], reps: int):
    """Same contract as the reference: rows shifted right by (reps-1) per earlier placeholder, tail dropped,
    placeholder rows duplicated; returns (tensor, final positions).  Implemented as one gather."""
    assert reps >= 1
    src, final = placeholder_row_map(ori.shape[0], r_pos, reps)
    idx = torch.as_tensor(src, device=ori.device, dtype=torch.long)
    ori[:] = ori[idx]
    return ori, final
