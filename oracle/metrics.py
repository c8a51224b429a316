"""Oracle: per-class intersection / union and mIoU (numpy, integer-exact).

Restates ``utils/metrics.py:6-19`` (``intersectionAndUnionGPU``: three
``torch.histc`` calls over [0, K-1] with K bins, after ``pred[target==255]=255``)
and the accumulation / mIoU post-processing of
``workflows/trainer/base_trainer.py:160-186`` and ``workflows/validator.py:85-115``.

``torch.histc(x, bins=K, min=0, max=K-1)`` on integer-valued floats is a plain
bincount of the values that fall in [0, K-1] (bin = floor(x*K/(K-1)), the top edge
folded into the last bin), so the restatement uses ``np.bincount``; results are
returned as float32 like the reference's.  Test infrastructure only.
"""

from __future__ import annotations

import numpy as np

IGNORE = 255


def confusion_matrix(pred, target, K, ignore_index=IGNORE):
    """int64 [K+1, K+1]; rows = target, cols = pred; index K collects values outside
    [0, K) (for pred this includes 255).  Pixels whose target == ignore_index are dropped."""
    pred = np.asarray(pred).astype(np.int64).ravel()
    target = np.asarray(target).astype(np.int64).ravel()
    keep = target != ignore_index
    p = pred[keep]
    t = target[keep]
    p = np.where((p >= 0) & (p < K), p, K)
    t = np.where((t >= 0) & (t < K), t, K)
    return np.bincount(t * (K + 1) + p, minlength=(K + 1) ** 2).reshape(K + 1, K + 1)


def intersection_and_union(pred, target, K, ignore_index=IGNORE):
    """(area_intersection, area_union) float32 [K]  (metrics.py:6-19).

    Also returns the mutated prediction (the reference overwrites the caller's
    tensor in place at :12).
    """
    pred = np.array(pred, dtype=np.int64, copy=True).ravel()
    target = np.asarray(target).astype(np.int64).ravel()
    pred[target == ignore_index] = ignore_index                            # :12

    def hist(x):
        x = x[(x >= 0) & (x <= K - 1)]
        return np.bincount(x, minlength=K).astype(np.float32)

    inter = hist(pred[pred == target])                                     # :13,15
    area_out = hist(pred)                                                  # :16
    area_tgt = hist(target)                                                # :17
    return inter, area_out + area_tgt - inter, pred


def iou_from_sums(intersection_sum, union_sum, synthia=False):
    """base_trainer.py:183-184 / validator.py:105-113."""
    iou = np.asarray(intersection_sum) / (np.asarray(union_sum) + 1e-10)
    miou = np.mean(iou)
    out = {'iou': iou, 'miou': miou}
    if synthia:
        out['miou_16'] = miou * 19 / 16
        iu_13 = iou.copy()
        iu_13[3:6] = 0
        out['miou_13'] = np.mean(iu_13) * 19 / 13
    return out
