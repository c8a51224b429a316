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
        """:29-34  (1 - v)^2 / sum with the reference's own TORCH float64 ops (``torch.tensor(...)``, ``**``, ``torch.sum``):
        numpy's summation order can differ in the last ulp, and a last-ulp difference in p can change a draw of
        ``np.random.choice`` (pinned on 20 seeds by tests/golden/copy_paste_probs.npz).  Classes forced to inf (SYNTHIA) get
        probability 0 instead of the reference's NaN; for finite values nothing differs from the reference."""
        probs = torch.tensor(np.asarray(self.class_value, dtype=np.float64))
        finite = torch.isfinite(probs)
        if bool(finite.all()):
            probs = (1 - probs) ** 2
        else:
            zero = torch.zeros_like(probs)
            probs = torch.where(finite, (1 - torch.where(finite, probs, zero)) ** 2, zero)
        probs = probs / torch.sum(probs)
        return probs.numpy()

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

    def run_batch(self, imgs, lbls, donor_imgs=None, donor_lbls=None, donor_index=None):
        """One launch for a whole batch already on the device.  imgs u8 [N,H,W,3], lbls u8 [N,H,W].  With explicit donors the
        donor of image i is donor_*[donor_index[i]]; without them the batch-level sampler draws one donor per image exactly as
        ``run_original`` would for the images in order (same global ``np.random`` stream), loads the donors from
        ``dataset_copy_from`` and uploads each distinct donor once (``DonorSampler``).  Returns (imgs, lbls, masks), in place."""
        if donor_imgs is None:
            if getattr(self, '_sampler', None) is None:
                self._sampler = DonorSampler(self)
            donor_imgs, donor_lbls, donor_index = self._sampler.sample(lbls.shape[0], tuple(lbls.shape[1:]))
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


class DonorSampler:
    """Batch-level donor sampler for the GPU copy-paste (VERDICT r1 missing #4).

    The reference pastes inside DataLoader workers, one image at a time (``base_dataset.py:99-124`` ->
    ``preprocessor.py:79-122``): per image one class draw (``random_select``, :93), one file draw (:95), one ``load_data``
    (:97).  On the device the paste is one launch per batch, so the donors of a batch are drawn up front -- the same two
    ``np.random`` calls per image in image order, hence the same donors as a sequential reference run with the same seed --
    loaded on a small thread pool, staged in pinned memory and uploaded into a device SLAB of ``cache`` donor slots; the
    kernel reads donor i of the batch from slot ``donor_index[i]``.  A donor that repeats (within the batch or across
    batches) is loaded and uploaded once; the least recently used slots are recycled.

    (``run_original``'s ``for _ in range(3)`` loop always stops after its first donor: the first pass marks every hard class
    as existing regardless of the donor's content, SURVEY.md A.4 -- one donor per image is the reference's behaviour.)"""

    def __init__(self, copy_paste, cache=64, workers=4):
        from concurrent.futures import ThreadPoolExecutor
        self.cp = copy_paste
        self.cache = int(cache)
        self._slot_of = {}                            # dataset index -> slab slot
        self._used = {}                               # slab slot -> last use tick
        self._tick = 0
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)))
        self._slab = self._pins = self._uploaded = None

    def draw(self, n):
        """Dataset indices of the donors of n images in order: the RNG calls of preprocessor.py:93-96, nothing else."""
        cp = self.cp
        out = []
        for _ in range(n):
            select_c = cp.random_select(cp.hard_classes)
            file_name = np.random.choice(cp.samples_with_class[select_c])
            out.append(cp.dataset_copy_from.get_file_to_idx(file_name))
        return out

    def _load(self, idx, hw):
        img_, lbl_, _path = self.cp.dataset_copy_from.load_data(idx)
        if tuple(img_.shape[:2]) != tuple(hw):                                  # :99-100
            img_, lbl_ = self.cp.resize(img_, lbl_, hw)
        return np.ascontiguousarray(img_, dtype=np.uint8), np.ascontiguousarray(lbl_, dtype=np.uint8)

    def sample(self, n, hw):
        """(donor slab imgs u8 [S,H,W,3], slab lbls u8 [S,H,W], donor_index i32 [n]) on the device for a batch of n images."""
        dev = self.cp.device
        h, w = int(hw[0]), int(hw[1])
        idxs = self.draw(n)
        distinct = list(dict.fromkeys(idxs))
        slots = max(self.cache, len(distinct))
        if self._slab is None or tuple(self._slab[1].shape[1:]) != (h, w) or self._slab[1].shape[0] < slots:
            self._slab = (torch.empty((slots, h, w, 3), dtype=torch.uint8, device=dev),
                          torch.empty((slots, h, w), dtype=torch.uint8, device=dev))
            self._slot_of, self._used = {}, {}
        missing = [i for i in distinct if i not in self._slot_of]
        if missing:
            total = self._slab[1].shape[0]
            free = [s for s in range(total) if s not in self._used]
            if len(free) < len(missing):              # recycle the least recently used slots that this batch does not need
                keep = {self._slot_of[i] for i in distinct if i in self._slot_of}
                victims = sorted((t, s) for s, t in self._used.items() if s not in keep)[:len(missing) - len(free)]
                gone = {s for _, s in victims}
                self._slot_of = {i: s for i, s in self._slot_of.items() if s not in gone}
                for s in gone:
                    del self._used[s]
                free += sorted(gone)
            loaded = list(self._pool.map(lambda i: self._load(i, (h, w)), missing))
            m = len(missing)
            if self._uploaded is not None:
                self._uploaded.synchronize()          # the previous batch's uploads still read the pinned staging buffers
            if self._pins is None or self._pins[1].shape[0] < m or tuple(self._pins[1].shape[1:]) != (h, w):
                self._pins = (torch.empty((max(m, 8), h, w, 3), dtype=torch.uint8).pin_memory(),
                              torch.empty((max(m, 8), h, w), dtype=torch.uint8).pin_memory())
            for k, (i, (img_, lbl_)) in enumerate(zip(missing, loaded)):
                s = free[k]
                self._pins[0][k].copy_(torch.from_numpy(img_))
                self._pins[1][k].copy_(torch.from_numpy(lbl_))
                self._slab[0][s].copy_(self._pins[0][k], non_blocking=True)
                self._slab[1][s].copy_(self._pins[1][k], non_blocking=True)
                self._slot_of[i] = s
            self._uploaded = torch.cuda.Event()
            self._uploaded.record(torch.cuda.current_stream(dev))
        self._tick += 1
        for i in distinct:
            self._used[self._slot_of[i]] = self._tick
        donor_index = torch.tensor([self._slot_of[i] for i in idxs], dtype=torch.int32).to(dev, non_blocking=True)
        return self._slab[0], self._slab[1], donor_index
