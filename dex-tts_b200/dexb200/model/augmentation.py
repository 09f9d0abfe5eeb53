"""``model.augmentation.Augment`` -- SpecAugment-style masking / circular shift of one (n_mels, T) feature map.

Training-side data augmentation of the reference (DEX-TTS/model/augmentation.py:9-73), imported by its data loader
(DEX-TTS/src/dataset.py:11,22) and therefore part of the import chain of ``main.py``.  CPU tensor work inside ``Dataset.__getitem__``;
nothing here touches the CUDA path.  Behaviour kept, including what looks accidental upstream:
  * the mask width is ``int(U[0, para))`` from ``numpy.random`` and its start ``random.randint(0, size - width)`` (inclusive), one
    (width, start) pair per mask, drawn in that order -- so a seeded run consumes the two generators exactly like the reference;
  * the shift point is ``int(numpy.random.uniform(T))`` -- ``T`` lands in ``low`` and ``high`` keeps its default 1.0, i.e. the point
    is drawn from (1, T] (augmentation.py:52);
  * ``aug_type`` is searched for 'T', then 'F', then 'S' and only the first hit is applied (augmentation.py:63-70);
  * masks are written in place into the caller's tensor; the result is returned through ``squeeze(1)`` (augmentation.py:72).
"""
import random

import numpy as np
import torch
from torch import nn


class Augment(nn.Module):
    def __init__(self, freq_mask_num=1, time_mask_num=1, freq_mask=False, time_mask=True):
        super().__init__()
        self.freq_mask_num, self.time_mask_num = freq_mask_num, time_mask_num
        self.freq_mask, self.time_mask = freq_mask, time_mask

    @staticmethod
    def _zero_bands(x, axis, count, max_width):
        size = x.shape[axis]
        for _ in range(count):
            width = int(np.random.uniform(low=0.0, high=max_width))
            start = random.randint(0, size - width)
            x.narrow(axis, start, width).zero_()
        return x

    def freq_mask_augment(self, x, freq_mask_para):
        return self._zero_bands(x, 0, self.freq_mask_num, freq_mask_para)

    def time_mask_augment(self, x, time_mask_para):
        return self._zero_bands(x, 1, self.time_mask_num, time_mask_para)

    def shift_augment(self, x):
        cut = int(np.random.uniform(x.shape[1]))
        return torch.roll(x, shifts=-cut, dims=1) if x.shape[1] else x

    def forward(self, x, aug_type, time_mask_para=27, freq_mask_para=30):
        if x.dim() != 2:
            x = x.unsqueeze(0)
        for tag, fn in (("T", lambda t: self.time_mask_augment(t, time_mask_para)),
                        ("F", lambda t: self.freq_mask_augment(t, freq_mask_para)),
                        ("S", self.shift_augment)):
            if tag in aug_type:
                x = fn(x)
                break
        return x.squeeze(1)
