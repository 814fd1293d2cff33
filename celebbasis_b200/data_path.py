"""Device-side data path: what FaceIdDatasetStyleGAN3.__getitem__ does per sample
(ldm/data/face_id.py:598-644), split into

  * the random DRAWS, made on the host with the same torch / numpy / random calls in the same order as the reference
    (RandomHorizontalFlip: torch.rand(1); ColorJitter.get_params: torch.randperm(4) + four torch.empty(1).uniform_;
    `np.random.randint(10)` for the never-taken dual-image branch; _add_bg: two np.random.uniform + up to two
    np.random.randint; random.choice for the caption) -- so a seeded run draws exactly what the reference draws;
  * the PIXEL work (flip, colour jitter, normalise, rescale + paste), which runs as cb_face_augment / cb_paste_resized on
    uint8 images already resident on the device instead of in 8 PIL worker processes.
"""
import random

import numpy as np
import torch

from . import lib as _lib
from . import ops

BRIGHTNESS, CONTRAST, SATURATION, HUE = (0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.01, 0.01)


def draw_trans_params(flip_p=0.5, jitter=True):
    """One `self.trans(img)` call (face_id.py:526-532): RandomHorizontalFlip then ColorJitter.get_params.
    Returns (iparams [5] int32: flip, op order; fparams [4] float32: brightness, contrast, saturation, hue)."""
    flip = int(bool(torch.rand(1) < flip_p))                                  # transforms.RandomHorizontalFlip.forward
    if jitter:
        order = torch.randperm(4).tolist()                                    # ColorJitter.get_params
        b = float(torch.empty(1).uniform_(*BRIGHTNESS))
        c = float(torch.empty(1).uniform_(*CONTRAST))
        s = float(torch.empty(1).uniform_(*SATURATION))
        h = float(torch.empty(1).uniform_(*HUE))
    else:                                                                     # split == 'dev': no jitter
        order, (b, c, s, h) = [-1, -1, -1, -1], (1.0, 1.0, 1.0, 0.0)
    return np.asarray([flip] + order, dtype=np.int32), np.asarray([b, c, s, h], dtype=np.float32)


def draw_add_bg(h, w, scale=(0.1, 1.0)):
    """_add_bg's draws (face_id.py:451-470): (rh, rw, pos_h, pos_w)."""
    rh = min(int(h * np.random.uniform(scale[0], scale[1])), h)
    rw = min(int(rh * np.random.uniform(0.9, 1.1)), w)
    pos_h = np.random.randint(h - rh) if h > rh else 0
    pos_w = np.random.randint(w - rw) if w > rw else 0
    return np.asarray([rh, rw, pos_h, pos_w], dtype=np.int32)


def identity_bg(h, w):
    return np.asarray([h, w, 0, 0], dtype=np.int32)


def face_augment(src_u8, iparams, fparams, out=None, c_off=0):
    """src_u8 (B,H,W,3) uint8 on the device; iparams (B,5) int32, fparams (B,4) float32 (device) -> fp32 (B,H,W,C)."""
    B, H, W, _ = src_u8.shape
    assert src_u8.dtype == torch.uint8 and src_u8.is_cuda and src_u8.is_contiguous()
    if out is None:
        out = torch.empty(B, H, W, 3, dtype=torch.float32, device=src_u8.device)
    ws = torch.empty(B, dtype=torch.float64, device=src_u8.device)
    _lib.check(_lib.load().cb_face_augment(ops._p(src_u8), ops._p(iparams), ops._p(fparams), ops._p(ws), ops._p(out), B, H, W,
                                           out.shape[-1], c_off, ops._st()), "cb_face_augment")
    return out


def paste_resized(faces, geo, c_off=0):
    """faces (B,H,W,C) fp32, geo (B,4) int32 {rh, rw, pos_h, pos_w} (device) -> image (B,H,W,3) fp32 (_add_bg)."""
    B, H, W, C = faces.shape
    out = torch.empty(B, H, W, 3, dtype=torch.float32, device=faces.device)
    _lib.check(_lib.load().cb_paste_resized(ops._p(faces), C, c_off, ops._p(geo), ops._p(out), B, H, W, ops._st()),
               "cb_paste_resized")
    return out


RAW_KEYS = ("image_u8", "aug_i", "aug_f", "aug_geo")


def is_raw_batch(batch):
    return isinstance(batch, dict) and "image_u8" in batch


def device_augment(batch, device):
    """A collated RAW batch (ldm.data.face_id mirror: uint8 images + the per-sample draws) -> the batch dict the reference's
    DataLoader yields: `image` (B,H,W,3) fp32 in [-1,1] after _add_bg, `image_ori.faces` (B,H,W,3k) fp32 (the jittered,
    un-pasted face stack), `image_ori.ids`, `image_ori.num_ids`, `caption`."""
    u8 = batch["image_u8"].to(device, non_blocking=True)                   # (B, k, H, W, 3) uint8
    B, k, H, W, _ = u8.shape
    ai = batch["aug_i"].to(device, non_blocking=True).int().contiguous()    # (B, k, 5)
    af = batch["aug_f"].to(device, non_blocking=True).float().contiguous()  # (B, k, 4)
    geo = batch["aug_geo"].to(device, non_blocking=True).int().contiguous()  # (B, 4)
    faces = torch.empty(B, H, W, 3 * k, dtype=torch.float32, device=device)
    for j in range(k):
        face_augment(u8[:, j].contiguous(), ai[:, j].contiguous(), af[:, j].contiguous(), out=faces, c_off=3 * j)
    image = paste_resized(faces, geo, c_off=0)
    out = {kk: v for kk, v in batch.items() if kk not in RAW_KEYS}
    out["image"] = image
    io = dict(batch["image_ori"])
    io["faces"] = faces
    out["image_ori"] = io
    return out


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
