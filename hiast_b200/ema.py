"""EMA teacher update in one multi-tensor launch -- mirror of ``utils/utils.py:115-123`` (``update_ema_model``).

    from hiast_b200.ema import update_ema_model
    update_ema_model(ema_model, model, gamma)        # same call as the reference's utils.update_ema_model

The reference runs five eager kernels per parameter tensor (two clones, two multiplies, one add; ~1500 launches per
training iteration for DeepLabv2-ResNet101).  Here the (teacher, student) parameter pointers are gathered once into a
device table and ``hiast_ema_update`` streams every parameter in one launch; buffers are copied by ``hiast_multi_copy``.
The arithmetic is the reference's (float32 ``k * gamma``, float32 ``q * (1 - gamma)``, float32 sum).  The update is in
place (the reference re-binds ``param_k.data`` to a fresh tensor; the values are identical).  No CPU path.
"""

from __future__ import annotations

import weakref

import numpy as np
import torch

from ._lib import HiastError, check, lib, ptr, stream_ptr

CHUNK_ELEMS = 1 << 15
CHUNK_BYTES = 1 << 18
# ema_model -> {id(model): entry}; entries hold a weak reference to ``model`` and die with either module, so a recycled
# ``id()`` can never hit a stale entry and dead models' parameters are not kept alive
_cache = weakref.WeakKeyDictionary()


class _Tables:
    def __init__(self, pairs, unit, chunk, device):
        segs, chunk_seg, chunk_off = [], [], []
        for s, (k, q) in enumerate(pairs):
            n = k.numel() if unit == 'elems' else k.numel() * k.element_size()
            segs.append((k.data_ptr(), q.data_ptr(), n))
            for off in range(0, n, chunk):
                chunk_seg.append(s)
                chunk_off.append(off)
        self.ptrs = [(k.data_ptr(), q.data_ptr()) for k, q in pairs]
        self.n_chunks = len(chunk_seg)
        self.chunk = chunk
        self.segs = torch.tensor(np.asarray(segs, dtype=np.int64).reshape(-1, 3), device=device)
        self.chunk_seg = torch.tensor(np.asarray(chunk_seg, dtype=np.int32), device=device)
        self.chunk_off = torch.tensor(np.asarray(chunk_off, dtype=np.int64), device=device)


def reset():
    """Forget the cached parameter lists / pointer tables (call after replacing Parameter objects in a model: the
    module trees are walked only on the first call for a given (ema_model, model) pair; re-allocated ``.data`` is
    detected automatically)."""
    for k in list(_cache.keys()):
        del _cache[k]


def _validate(params, buffers):
    for k, q in params:
        if not (k.is_cuda and q.is_cuda):
            raise HiastError('update_ema_model: parameters must live on a CUDA device (there is no CPU path)')
        if k.dtype != torch.float32 or q.dtype != torch.float32:
            raise NotImplementedError('update_ema_model: float32 parameters only (got %s / %s)' % (k.dtype, q.dtype))
        if k.shape != q.shape or not (k.is_contiguous() and q.is_contiguous()):
            raise HiastError('update_ema_model: parameter pairs must be contiguous and of equal shape')
    for k, q in buffers:
        if k.dtype != q.dtype or k.shape != q.shape or not (k.is_cuda and q.is_cuda and k.is_contiguous() and q.is_contiguous()):
            raise HiastError('update_ema_model: buffer pairs must be contiguous CUDA tensors of equal dtype and shape')


def update_ema_model(ema_model, model, gamma):
    """utils/utils.py:115-123.  Returns ``ema_model`` like the reference."""
    per_teacher = _cache.get(ema_model)
    if per_teacher is None:
        per_teacher = _cache[ema_model] = {}
    ent = per_teacher.get(id(model))
    if ent is not None:
        # same objects as when the trees were walked?  (weak reference still alive and pointing at THIS model; the first
        # Parameter of both trees unchanged -- replacing other Parameter objects needs reset(), see its docstring)
        first_q, first_k = next(model.parameters(), None), next(ema_model.parameters(), None)
        stale = ent['model']() is not model or (ent['params'] and (ent['params'][0][1] is not first_q or ent['params'][0][0] is not first_k))
        if stale:
            ent = None
    if ent is None:
        # the module trees are walked once; later calls only re-read the data pointers of the cached tensors
        for dead in [k for k, e in per_teacher.items() if e['model']() is None]:
            del per_teacher[dead]
        ent = per_teacher[id(model)] = {'model': weakref.ref(model),
                                        'params': [(k, q) for q, k in zip(model.parameters(), ema_model.parameters())],
                                        'buffers': [(k, q) for q, k in zip(model.buffers(), ema_model.buffers())], 'ptrs': None}
    params, buffers = ent['params'], ent['buffers']
    if not params and not buffers:
        return ema_model
    ptrs = [t.data_ptr() for pair in params for t in pair] + [t.data_ptr() for pair in buffers for t in pair]
    if ent['ptrs'] != ptrs:
        _validate(params, buffers)
        device = (params or buffers)[0][0].device
        ent['tp'] = _Tables(params, 'elems', CHUNK_ELEMS, device)
        ent['tb'] = _Tables([(k, q) for k, q in buffers if k.numel()], 'bytes', CHUNK_BYTES, device)
        ent['ptrs'] = ptrs
        ent['device'] = device
    tp, tb = ent['tp'], ent['tb']
    st = stream_ptr(ent['device'])
    g32, omg32 = float(np.float32(gamma)), float(np.float32(1.0 - gamma))
    if tp.n_chunks:
        check(lib().hiast_ema_update(ptr(tp.segs), ptr(tp.chunk_seg), ptr(tp.chunk_off), tp.n_chunks, tp.chunk, g32, omg32, st),
              'hiast_ema_update')
    if tb.n_chunks:
        check(lib().hiast_multi_copy(ptr(tb.segs), ptr(tb.chunk_seg), ptr(tb.chunk_off), tb.n_chunks, tb.chunk, st),
              'hiast_multi_copy')
    return ema_model
