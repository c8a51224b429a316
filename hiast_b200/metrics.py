"""Evaluation metric on the shared-memory privatised confusion-matrix kernel.

Mirrors ``utils/metrics.py:6-19`` (``intersectionAndUnionGPU``) and the accumulation / mIoU code of
``workflows/trainer/base_trainer.py:160-186`` and ``workflows/validator.py:85-115`` (reference,
/root/reference/code).  The reference derives intersection / union from three ``torch.histc`` calls
after compacting and casting; here one kernel builds the (K+1)x(K+1) int64 confusion matrix, from
which the same float32 vectors are read off (diag / row sums / column sums).
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import HiastError

IGNORE = 255


def _flat(t, name):
    if not t.is_cuda:
        raise HiastError('%s must be a CUDA tensor (there is no CPU path)' % name)
    if t.dtype not in (torch.uint8, torch.int64):
        raise HiastError('%s must be uint8 or int64' % name)
    return t


def confusion_matrix(output, target, K, ignore_index=IGNORE, cm=None):
    """int64 [K+1,K+1]; rows = target, cols = pred, index K = value outside [0,K).  Does not mutate."""
    output, target = _flat(output, 'output'), _flat(target, 'target')
    if output.dtype != target.dtype:
        output, target = output.long(), target.long()
    return ops.confusion_matrix(output.contiguous(), target.contiguous(), K, ignore_index, cm)


def intersectionAndUnionGPU(output, target, K, ignore_index=IGNORE):
    """metrics.py:6-19: (area_intersection, area_union), float32 [K] on the input device.

    Like the reference it overwrites ``output[target == ignore_index] = ignore_index`` in the caller's
    tensor (:12) when ``output`` is a contiguous int64 / uint8 tensor of the same dtype as ``target``.
    """
    assert output.dim() in [1, 2, 3]                                       # :8
    assert output.shape == target.shape                                    # :9
    output, target = _flat(output, 'output'), _flat(target, 'target')
    same = output.dtype == target.dtype and output.is_contiguous() and target.is_contiguous()
    if same:
        cm = ops.confusion_matrix(output, target, K, ignore_index, mutate_pred=True)
    else:
        cm = ops.confusion_matrix(output.long().contiguous(), target.long().contiguous(), K, ignore_index)
        output[target == ignore_index] = ignore_index
    return ops.iou_from_confusion(cm, K)


class ConfusionMeter:
    """Running evaluation state (base_trainer.py:162-184, validator.py:85-113).

    Keeps both the exact int64 confusion matrix and the reference's float32 running sums
    (``intersection_sum += intersection`` in fp32 loses integer exactness past 2^24; it is kept only
    to reproduce the reference's printed numbers).
    """

    def __init__(self, num_classes, device='cuda', ignore_index=IGNORE):
        self.K = num_classes
        self.ignore_index = ignore_index
        self.cm = torch.zeros((num_classes + 1, num_classes + 1), dtype=torch.int64, device=device)
        self.intersection_sum = torch.zeros(num_classes, dtype=torch.float32, device=device)
        self.union_sum = torch.zeros(num_classes, dtype=torch.float32, device=device)

    def update(self, lbl_pred, lbl):
        step = confusion_matrix(lbl_pred, lbl, self.K, self.ignore_index)
        inter, union = ops.iou_from_confusion(step, self.K)
        self.cm += step
        self.intersection_sum += inter
        self.union_sum += union
        return inter, union

    def update_from_logits(self, logits, lbl):
        """argmax(dim=1) fused into the bincount (base_trainer.py:173-175)."""
        step = ops.confusion_from_logits(logits.contiguous(), lbl.contiguous(), self.K, self.ignore_index)
        inter, union = ops.iou_from_confusion(step, self.K)
        self.cm += step
        self.intersection_sum += inter
        self.union_sum += union
        return inter, union

    def all_reduce(self):
        """base_trainer.py:180-181 (+ the exact matrix)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.cm)
            dist.all_reduce(self.intersection_sum)
            dist.all_reduce(self.union_sum)

    def result(self, synthia=False, exact=False):
        """iou / miou as the reference computes them (:183-184; validator.py:105-113)."""
        if exact:
            inter, union = ops.iou_from_confusion(self.cm, self.K)
        else:
            inter, union = self.intersection_sum, self.union_sum
        iou = inter.cpu().numpy() / (union.cpu().numpy() + 1e-10)
        miou = np.mean(iou)
        out = {'iou': iou, 'miou': miou}
        if synthia:
            out['miou_16'] = miou * 19 / 16
            iu_13 = iou.copy()
            iu_13[3:6] = 0
            out['miou_13'] = np.mean(iu_13) * 19 / 13
        return out
