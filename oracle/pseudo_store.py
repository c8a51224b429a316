"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reader side of the on-disk pseudo-label outputs.

Follows sseg/datasets/loader/base_dataset.py (reference, /root/reference/code): ``stat_samples_with_class`` :61-77 and
the pseudo-label branch of ``load_data`` :158-178.  Pinned against tests/golden/pseudo_store.npz, produced by running the
unmodified reference functions (tests/golden/make_golden.py).  ``cv2.resize(..., INTER_NEAREST)`` is OpenCV (not under
/root/reference; installed 4.13): its published rule ``sx = min(floor(x * (1 / (dst / src))), src - 1)`` in double precision
is restated in numpy and checked against the fixture and against cv2 itself.

Only tests/ may import this module.
"""

from __future__ import annotations

import numpy as np


def stat_samples_with_class(samples_with_class_and_n, num_classes):
    """:61-77.  Input: the parsed samples_with_class.json ({'class': [[path, pixels], ...]}); per class the file names
    sorted by pixel count (stable), the lowest round(10 %) dropped."""
    data = {int(k): v for k, v in samples_with_class_and_n.items()}
    out = {}
    for c in range(num_classes):
        names = [f.split('/')[-1] for f, _ in sorted(data[c], key=lambda item: item[1])]
        out[c] = names[round(len(names) * 0.1):]
    return out


def resize_nearest(lbl, size):
    """cv2.resize(lbl, (W, H), interpolation=cv2.INTER_NEAREST) for a 2-D uint8 array; size = (H, W)."""
    hs, ws = lbl.shape
    hd, wd = size
    ifx = 1.0 / (float(wd) / float(ws))
    ify = 1.0 / (float(hd) / float(hs))
    sx = np.minimum(np.floor(np.arange(wd) * ifx).astype(np.int64), ws - 1)
    sy = np.minimum(np.floor(np.arange(hd) * ify).astype(np.int64), hs - 1)
    return lbl[sy][:, sx]
