"""CPU oracle of the EMA teacher update (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Reference: ``utils/utils.py:115-123`` (``update_ema_model``), called once per training iteration by
``workflows/trainer/consistency_self_training_trainer.py`` after ``optimizer.step()``:

    param_k = param_k * gamma + param_q * (1 - gamma)        for every parameter pair   (:117-119)
    buffer_k = buffer_q                                       for every buffer pair      (:120-121)

``gamma`` is a Python float and the parameters are float32 tensors, so torch multiplies by the float32 roundings of
``gamma`` and of the DOUBLE ``1 - gamma`` and rounds after every one of the three operations (no fma contraction in
eager mode).
"""

import numpy as np


def ema_update(params_k, params_q, gamma):
    """params_*: lists of float32 numpy arrays.  Returns the new teacher parameters (:117-119)."""
    g = np.float32(gamma)
    omg = np.float32(1.0 - gamma)
    out = []
    for k, q in zip(params_k, params_q):
        k = np.asarray(k, dtype=np.float32)
        q = np.asarray(q, dtype=np.float32)
        out.append((k * g).astype(np.float32) + (q * omg).astype(np.float32))
    return out


def copy_buffers(buffers_q):
    """:120-121"""
    return [np.array(b, copy=True) for b in buffers_q]
