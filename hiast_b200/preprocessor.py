"""Hard-aware pseudo-label augmentation (CopyPaste) on the masked-gather CUDA kernel.

Mirrors ``sseg/datasets/preprocessor.py`` (reference, /root/reference/code): ``CopyPaste.__init__``
:14-27, ``calculate_class_probs`` :29-34, ``get_hard_classes`` :36-44, ``random_select`` :70-77,
``run`` :64-68, ``run_original`` :79-122.  Host logic (hard-class ranking, sampling probabilities,
donor choice from the global ``np.random`` stream) is identical call for call, so a seeded run picks
the same donors as the reference; the 14 compares + masked stores + 2 fancy-index copies per donor
(:103-112) are one kernel pass.

The reference runs this inside DataLoader worker processes on numpy arrays; a CUDA context cannot
live there, so this class must be called from the main process: ``run(img, lbl)`` accepts numpy
arrays (copied to the device and back; drop-in but bounded by PCIe) or CUDA uint8 tensors (stay on
the device), and ``run_batch`` pastes a whole batch with one launch.

SYNTHIA: the reference sets class_value[9,14,16] = inf, which turns its sampling probabilities into
NaN and makes ``np.random.choice`` raise.  Here those classes get probability 0.
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops
from .registry import PREPROCESSOR


@PREPROCESSOR.register('CopyPaste')
class CopyPaste:

    def __init__(self, cfg, dataset_copy_from, init_class_value, device='cuda'):
        self.cfg = cfg
        self.dataset_copy_from = dataset_copy_from
        self.device = torch.device(device)
        if self.cfg.dataset.source.type == 'SYNTHIA':                      # :18-21
            self.ignored_classes = [9, 14, 16]
        else:
            self.ignored_classes = None
        self.class_value, self.hard_classes = self.get_hard_classes(init_class_value)
        self.samples_with_class = self.dataset_copy_from.get_samples_with_class()
        self.class_probs = self.calculate_class_probs()

    def calculate_class_probs(self):
        """:29-34  (1 - v)^2 / sum; classes forced to inf get probability 0 instead of NaN."""
        v = np.asarray(self.class_value, dtype=np.float64)
        finite = np.isfinite(v)
        p = np.where(finite, (1 - np.where(finite, v, 0.0)) ** 2, 0.0)
        return p / np.sum(p)

    def get_hard_classes(self, class_value):
        """:36-44"""
        if self.ignored_classes is not None:
            for c in self.ignored_classes:
                class_value[c] = np.inf
        hard_classes = np.argsort(class_value)[:self.cfg.preprocessor.copy_paste.selected_num_classes]
        return class_value, hard_classes

    def random_select(self, selected_classes):
        """:70-77"""
        while True:
            select_c = np.random.choice([i for i in range(self.cfg.dataset.num_classes)], size=1, replace=False,
                                        p=self.class_probs)[0]
            if select_c in selected_classes:
                break
        return select_c

    def run(self, img, lbl):
        """:64-68"""
        if self.cfg.preprocessor.copy_paste.mode == 'original':
            return self.run_original(img, lbl)
        return NotImplementedError

    # ------------------------------------------------------------------ donors
    def _choose_donor(self, selected_classes):
        select_c = self.random_select(selected_classes)                    # :93
        file_name = np.random.choice(self.samples_with_class[select_c])    # :95
        tmp_idx = self.dataset_copy_from.get_file_to_idx(file_name)        # :96
        return self.dataset_copy_from.load_data(tmp_idx)                   # :97

    def _to_dev(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device, non_blocking=True).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=True)

    def run_original(self, img, lbl):
        """:79-122.  img uint8 [H,W,3], lbl uint8 [H,W]; returns (img, lbl, copy_paste_mask)."""
        as_numpy = not isinstance(img, torch.Tensor)
        d_img = self._to_dev(img).unsqueeze(0)
        d_lbl = self._to_dev(lbl).unsqueeze(0)
        if d_img.dtype != torch.uint8 or d_lbl.dtype != torch.uint8:
            raise TypeError('CopyPaste expects uint8 image and label arrays')
        mask = torch.full_like(d_lbl, 255)                                 # :87
        selected_classes = self.hard_classes
        exist_classes = []
        for _ in range(3):
            img_, lbl_, _path = self._choose_donor(selected_classes)
            if tuple(img.shape) != tuple(img_.shape):                      # :99-100
                img_, lbl_ = self.resize(img_, lbl_, lbl.shape)
            for c in self.hard_classes:                                    # :104-106
                if c in selected_classes and c not in exist_classes:
                    exist_classes.append(c)
            ops.copy_paste(d_img, d_lbl, mask, self._to_dev(img_).unsqueeze(0), self._to_dev(lbl_).unsqueeze(0),
                           self.hard_classes)
            non_exist_classes = [c for c in self.hard_classes if c not in exist_classes]
            if len(exist_classes) >= len(self.hard_classes) * 0.5:         # :117
                break
            selected_classes = non_exist_classes
        if as_numpy:
            out_img, out_lbl, out_mask = d_img[0].cpu().numpy(), d_lbl[0].cpu().numpy(), mask[0].cpu().numpy()
            img[...] = out_img                                             # the reference edits its inputs in place
            lbl[...] = out_lbl
            return img, lbl, out_mask
        return d_img[0], d_lbl[0], mask[0]

    def run_batch(self, imgs, lbls, donor_imgs, donor_lbls, donor_index=None):
        """One launch for a whole batch already on the device.  imgs u8 [N,H,W,3], lbls u8 [N,H,W]; the
        donor of image i is donor_*[donor_index[i]].  Returns (imgs, lbls, copy_paste_masks), in place."""
        masks = torch.full_like(lbls, 255)
        ops.copy_paste(imgs, lbls, masks, donor_imgs, donor_lbls, self.hard_classes, donor_index)
        return imgs, lbls, masks

    def resize(self, img, lbl, target_shape):
        """:46-52 (host OpenCV, only when donor and target shapes differ)."""
        import cv2
        if len(target_shape) > 2:
            target_shape = (target_shape[0], target_shape[1])
        img = cv2.resize(np.asarray(img), tuple(target_shape[::-1]), interpolation=cv2.INTER_LINEAR)
        lbl = cv2.resize(np.asarray(lbl), tuple(target_shape[::-1]), interpolation=cv2.INTER_NEAREST)
        return img, lbl
