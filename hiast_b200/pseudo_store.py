"""Reader side of the on-disk pseudo-label outputs.

Mirrors the parts of ``sseg/datasets/loader/base_dataset.py`` (reference, /root/reference/code) that consume what the
pseudo-label generators write:

* ``stat_samples_with_class`` :61-77 -- ``samples_with_class.json`` -> per class the image names sorted by pixel count with
  the lowest 10 % dropped (the donor pool of CopyPaste, preprocessor.py:26);
* the pseudo-label branch of ``load_data`` :158-178 -- ``{stem}_pseudo_label.png`` read with PIL and brought to the image
  size with ``cv2.resize(..., INTER_NEAREST)``.

The reference does the resize per sample in DataLoader workers; ``load_pseudo_labels`` decodes a batch of files on the
host (the files are the PNGs written by ``hiast_png_encode`` or by the reference) and resizes them with ONE launch of
``hiast_resize_nearest_u8`` on the device, where the copy-paste and loss kernels take them from.  The datasets themselves
(image decoding, augmentation) stay outside this package.
"""

from __future__ import annotations

import json
import os
import os.path as osp
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops


def stat_samples_with_class(data_root, num_classes):
    """:61-77"""
    with open(osp.join(data_root, 'samples_with_class.json'), 'r') as of:
        samples_with_class_and_n = json.load(of)
        samples_with_class_and_n = {int(k): v for k, v in samples_with_class_and_n.items()}
    samples_with_class = {}
    for c in range(num_classes):
        ranked = sorted(samples_with_class_and_n[c], key=lambda item: item[1])
        names = [file.split('/')[-1] for file, _pixels in ranked]
        samples_with_class[c] = names[round(len(names) * 0.1):]             # filter samples with too small pixels
    return samples_with_class


def pseudo_label_path(pseudo_dir, img_path):
    """:163-165"""
    return os.path.join(pseudo_dir, os.path.splitext(os.path.basename(img_path))[0] + '_pseudo_label.png')


def read_pseudo_label(pseudo_dir, img_path):
    """:168  ``np.array(Image.open(lbl_path), dtype=np.uint8)`` at the stored size."""
    from PIL import Image
    return np.array(Image.open(pseudo_label_path(pseudo_dir, img_path)), dtype=np.uint8)


def load_pseudo_labels(pseudo_dir, img_paths, size, device='cuda', workers=4):
    """Pseudo-labels of a batch of images as one uint8 CUDA tensor [N, size[0], size[1]] (:168 + :176 for every
    image).  Files are decoded on host threads; maps stored at another size are resized on the device."""
    if len(img_paths) == 0:
        return torch.empty((0, int(size[0]), int(size[1])), dtype=torch.uint8, device=device)
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        maps = list(pool.map(lambda p: read_pseudo_label(pseudo_dir, p), img_paths))
    size = (int(size[0]), int(size[1]))
    out = torch.empty((len(maps),) + size, dtype=torch.uint8, device=device)
    by_shape = {}
    for i, m in enumerate(maps):
        by_shape.setdefault(m.shape, []).append(i)
    for shape, idx in by_shape.items():
        stack = torch.from_numpy(np.stack([maps[i] for i in idx])).to(device, non_blocking=True)
        res = stack if tuple(shape) == size else ops.resize_nearest_u8(stack, size)
        out[torch.as_tensor(idx, device=device)] = res
    return out
