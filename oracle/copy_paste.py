"""Oracle: hard-aware pseudo-label augmentation (CopyPaste), numpy.

Restates ``sseg/datasets/preprocessor.py`` (reference, /root/reference/code):
``calculate_class_probs`` :29-34, ``get_hard_classes`` :36-44, ``random_select``
:70-77 and ``run_original`` :79-122.  The donor choice consumes the global
``np.random`` stream exactly like the reference (one or more
``np.random.choice(C, p)`` draws until a hard class comes up, then one
``np.random.choice(files)``), so seeding ``np.random.seed`` pins it.

SYNTHIA (classes 9, 14, 16 forced to +inf) makes the reference's sampling
probabilities NaN and ``np.random.choice`` raise; with
``nan_to_zero=True`` the oracle defines p_c = 0 for those classes instead (the
behaviour the build documents for BASELINE config 4).  Test infrastructure only.
"""

from __future__ import annotations

import numpy as np


def hard_classes(class_value, selected_num_classes, ignored_classes=None):
    """(class_value, hard) -- :36-44.  class_value is modified in place like the reference."""
    if ignored_classes is not None:
        for c in ignored_classes:
            class_value[c] = np.inf
    hard = np.argsort(class_value)[:selected_num_classes]
    return class_value, hard


def class_probs(class_value, nan_to_zero=False):
    """(1 - v)^2 / sum  (:29-34).  The reference computes this with TORCH float64 ops (``torch.tensor(class_value)``,
    ``**``, ``torch.sum``), whose summation order need not be numpy's; a last-ulp difference in p can change a draw of
    ``np.random.choice``, so the same ops are used here (pinned by tests/golden/copy_paste_probs.npz, 20 seeds)."""
    import torch
    probs = torch.tensor(np.asarray(class_value, dtype=np.float64))
    if nan_to_zero:
        finite = torch.isfinite(probs)
        probs = torch.where(finite, (1 - torch.where(finite, probs, torch.zeros_like(probs))) ** 2, torch.zeros_like(probs))
    else:
        probs = (1 - probs) ** 2
    probs = probs / torch.sum(probs)
    return probs.numpy()


def random_select(num_classes, probs, selected):
    """:70-77 -- rejection-sample a class from ``probs`` until it is in ``selected``."""
    while True:
        c = np.random.choice([i for i in range(num_classes)], size=1, replace=False, p=probs)[0]
        if c in selected:
            return c


def paste(img, lbl, cp_mask, donor_img, donor_lbl, hard):
    """One donor: the body of the loop at :102-112, in place on img / lbl / cp_mask."""
    sel = np.isin(donor_lbl, np.asarray(hard))
    cp_mask[sel] = donor_lbl[sel]
    img[sel] = donor_img[sel]
    lbl[sel] = donor_lbl[sel]
    return sel


def run_original(img, lbl, hard, probs, samples_with_class, load_donor, num_classes):
    """:79-122.  ``load_donor(file_name) -> (img_, lbl_)`` already at img's shape."""
    cp_mask = np.ones_like(lbl, dtype=np.uint8) * 255
    selected = hard
    exist = []
    donors = []
    for _ in range(3):
        c = random_select(num_classes, probs, selected)
        file_name = np.random.choice(samples_with_class[c])
        donors.append(file_name)
        d_img, d_lbl = load_donor(file_name)
        for k in hard:
            if k in selected and k not in exist:
                exist.append(k)
        paste(img, lbl, cp_mask, d_img, d_lbl, hard)
        missing = [k for k in hard if k not in exist]
        if len(exist) >= len(hard) * 0.5:
            break
        selected = missing
    return img, lbl, cp_mask, donors
