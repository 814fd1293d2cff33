"""Host mirror of ldm/data/face_id.py (FaceIdDatasetStyleGAN3 :471-731, FaceIdDatasetOneShot :751-757): same constructor
keywords, pickle format (a list of image paths, identity = file stem, gen_pickle.py), sample indexing, caption templates and
-- call for call -- the same random draws, but `__getitem__` returns the RAW sample: the uint8 image(s) plus the drawn
augmentation parameters.  The pixel work (flip, ColorJitter, normalise, _add_bg rescale + paste) runs on the device in
celebbasis_b200.data_path.device_augment (cb_face_augment / cb_paste_resized) once the collated batch has been copied
there, instead of in 8 PIL worker processes that cannot keep up with a ~16 ms training step."""
import os
import pickle
import random
import re

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from celebbasis_b200 import data_path


def _templates():
    """The 81 caption templates of face_id.py:17-99 (Textual Inversion's imagenet_templates_small, then the same list with
    'a photo of' replaced by 'an illustration of' / 'a depiction of'), in the reference's order (random.choice indexes it)."""
    tails = ["a {}", None, None, None, "a clean {}", "a dirty {}", None, "my {}", "the cool {}", None, None, None, "the {}",
             None, "one {}", None, None, "the clean {}", None, "a nice {}", None, "the nice {}", "the small {}",
             "the weird {}", "the large {}", "a cool {}", "a small {}"]
    fixed = {1: "a rendering of a {}", 2: "a cropped photo of the {}", 3: "the photo of a {}", 6: "a dark photo of the {}",
             9: "a close-up photo of a {}", 10: "a bright photo of the {}", 11: "a cropped photo of a {}",
             13: "a good photo of the {}", 15: "a close-up photo of the {}", 16: "a rendition of the {}",
             18: "a rendition of a {}", 20: "a good photo of a {}"}
    out = []
    for lead in ("a photo of", "an illustration of", "a depiction of"):
        for i, t in enumerate(tails):
            out.append(fixed[i] if t is None else f"{lead} {t}")
    return out


imagenet_templates_smallest = ['a photo of a {}']
imagenet_templates_small = _templates()
imagenet_dual_templates_small = [t.replace("{}", "{} and a {}") for t in imagenet_templates_small[:27]]
per_img_token_list = ['sks', 'ks', 'ata', 'tre', 'ry', 'bop', 'rn', '&', '*', '`']
reg_token_list = ['face']


class FaceIdDatasetStyleGAN3(Dataset):
    def __init__(self, pickle_path='/gavin/datasets/stylegan/stylegan3-r-ffhq-1024x1024_ffhq.pickle', num_ids=10,
                 specific_ids=None, image_size=512, repeats=100, flip_p=0.5, split="train", diff_cnt=32, **kwargs):
        super().__init__()
        images_per_id, reg_ids, reg_images_per_id, reg_repeats = 1, 0, 1, 0          # face_id.py:484-488 (hard-coded)
        if isinstance(specific_ids, str):
            if re.match(r'\d+-\d+', specific_ids) is None:
                raise ValueError('Specific_ids not supported.')
            lo, hi = [int(x) for x in specific_ids.split('-')]
            specific_ids = list(np.arange(lo, hi))
        self.pickle_path, self.num_ids = pickle_path, num_ids
        self.specific_ids = None if specific_ids is None else [int(s) for s in specific_ids]
        self.images_per_id, self.reg_ids, self.reg_images_per_id = images_per_id, reg_ids, reg_images_per_id
        self.diff_cnt = diff_cnt
        self.img_dict, self.img_list = {}, []
        self._load_from_pickle()
        self.repeats, self.reg_repeats, self.split = repeats, reg_repeats, split
        self.img_list = self.img_list[:num_ids * images_per_id] * repeats + self.img_list[num_ids * images_per_id:] * reg_repeats
        self.num_images = len(self.img_list)
        self.num_train = num_ids * images_per_id * repeats
        self.num_reg = reg_ids * reg_images_per_id * reg_repeats
        self._length = self.num_images
        self.image_size, self.flip_p = image_size, flip_p
        self._cache = {}
        print('[FaceIdDataset] loaded from %s. (id*max_img=%d*%d, train%d+reg%d=total%d)' % (
            pickle_path, num_ids, images_per_id, self.num_train, self.num_reg, self.num_images))

    def _load_from_pickle(self):
        with open(self.pickle_path, "rb") as handle:
            pickle_list = pickle.load(handle)
        pickle_dict = {}
        for img in pickle_list:
            pickle_dict.setdefault(os.path.basename(img).split('.')[0], []).append(img)
        walk_id = use_id = 0
        last = list(pickle_dict.keys())[-1]
        for id_, images in pickle_dict.items():
            if len(self.img_dict) >= self.num_ids:
                break
            if len(images) >= self.images_per_id:
                if self.specific_ids is not None and walk_id not in self.specific_ids:
                    walk_id += 1
                    continue
                self.img_dict[use_id] = {'id': id_, 'images': images[:self.images_per_id]}
                self.img_list += images[:self.images_per_id]
                walk_id += 1
                use_id += 1
            if id_ == last and use_id < self.num_ids:
                raise ValueError('Reach last. Not enough images for num_ids=%d, only %d.' % (self.num_ids, use_id))

    def __len__(self):
        return self._length

    def _get_id_and_img_idx(self, i):
        if i < self.num_train:
            i %= (self.num_ids * self.images_per_id)
            return i // self.images_per_id, i % self.images_per_id
        i -= self.num_train
        i %= (self.reg_ids * 1)
        return i // self.reg_images_per_id + self.num_ids, i % self.reg_images_per_id

    def _load_u8(self, path):
        """Image.open(...).convert('RGB') + transforms.Resize(image_size) (PIL bilinear, shorter edge), cached as uint8 HWC."""
        t = self._cache.get(path)
        if t is None:
            img = Image.open(path).convert('RGB')
            w, h = img.size
            if min(w, h) != self.image_size:
                s = self.image_size / min(w, h)
                img = img.resize((max(self.image_size, int(round(w * s))) if w > h else self.image_size,
                                  max(self.image_size, int(round(h * s))) if h > w else self.image_size), Image.BILINEAR)
            t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())
            self._cache[path] = t
        return t

    def _trans(self, path):
        ip, fp = data_path.draw_trans_params(self.flip_p, jitter=self.split != 'dev')
        return self._load_u8(path), ip, fp

    def __getitem__(self, i):
        id_idx, img_idx = self._get_id_and_img_idx(i)
        first = self._trans(self.img_dict[id_idx]['images'][img_idx])
        others, ids2 = [], []
        assert self.diff_cnt < self.num_ids
        for _ in range(self.diff_cnt):                                            # _get_diff_id_multi
            id2, im2 = id_idx, img_idx
            while id2 == id_idx:
                id2, im2 = self._get_id_and_img_idx(np.random.randint(self.num_train))
            others.append(self._trans(self.img_dict[id2]['images'][im2]))
            ids2.append(id2)
        ids3 = []
        for id3 in [id_idx] + ids2:                                               # _get_aug2_id_multi
            im3 = np.random.randint(self.images_per_id)
            others.append(self._trans(self.img_dict[id3]['images'][im3]))
            ids3.append(id3)
        np.random.randint(10)                                                     # dual_img is always False (:610)
        h, w = first[0].shape[:2]
        geo = data_path.draw_add_bg(h, w) if self.split != 'dev' else data_path.identity_bg(h, w)
        text = random.choice(imagenet_templates_small).format('face of %s person' % per_img_token_list[0])
        faces = [first] + others
        return {"image_u8": torch.stack([f[0] for f in faces], 0),
                "aug_i": torch.from_numpy(np.stack([f[1] for f in faces], 0)),
                "aug_f": torch.from_numpy(np.stack([f[2] for f in faces], 0)),
                "aug_geo": torch.from_numpy(geo),
                "image_ori": {"ids": torch.tensor([id_idx] + ids2 + ids3), "num_ids": 1},
                "caption": text, "id_idx": id_idx, "img_idx": img_idx}


class FaceIdDatasetOneShot(FaceIdDatasetStyleGAN3):
    def __init__(self, pickle_path: str, **kwargs):
        super().__init__(pickle_path, **kwargs)


class FaceIdDatasetE4T(FaceIdDatasetStyleGAN3):
    pass


class FaceIdDatasetNobody(FaceIdDatasetStyleGAN3):
    pass
