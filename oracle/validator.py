"""TEST INFRASTRUCTURE (oracle) -- torch restatement of the validator's multi-scale / flip prediction.

Follows workflows/validator.py:34-55 (``get_multi_scale_and_flip_logits``) and :92-93 (``argmax``) of the reference
(/root/reference/code).  Runs on whatever device the inputs live on: on CPU it is pinned against
tests/golden/validator.npz (the unmodified reference method run unbound on a ``SimpleNamespace``); on CUDA it IS the
reference's own PyTorch path (ATen softmax / interpolate / flip / add / argmax), which the kernels must match bit for bit.

Only tests/ may import this module.
"""

from __future__ import annotations

import torch
from torch.nn import functional as F


def multi_scale_and_flip(model, imgs, resize_sizes, is_flip, is_softmax=True):
    """:34-55.  ``model(x)`` returns {'logits': [B,C,h,w]} at the size of x."""
    pred_result_list = []
    if is_softmax:
        pred_fun = lambda x: F.softmax(model(x)['logits'], dim=1)        # noqa: E731  (:37)
    else:
        pred_fun = lambda x: model(x)['logits']                           # noqa: E731  (:39)
    for size in resize_sizes:
        assert len(size) == 2 and size[0] <= size[1]                      # :42-43
        tmp_imgs = F.interpolate(imgs, size, mode='bilinear', align_corners=True)
        pred_result = pred_fun(tmp_imgs)
        if is_flip:                                                       # :48-50
            flip_logits = pred_fun(torch.flip(tmp_imgs, dims=[3]))
            pred_result += torch.flip(flip_logits, dims=[3])
        pred_result = F.interpolate(pred_result, imgs.size()[2:], mode='bilinear', align_corners=True)
        pred_result_list.append(pred_result)
    return sum(pred_result_list)


def predict_labels(model, imgs, resize_sizes, is_flip):
    """:92-93"""
    return multi_scale_and_flip(model, imgs, resize_sizes, is_flip).argmax(dim=1)
