"""Functional wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function launches asynchronously on torch's current stream and never syncs with the host.
Shapes / dtypes follow include/hiast_b200.h.  No CPU path exists: CPU tensors raise.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import IGNORE, KEY_ONE, REGION, TERM_CE, TERM_CST, TERM_ENT, TERM_KLD, check, lib, ptr, require_cuda, stream_ptr

__all__ = [
    'ias_key_lo', 'ias_num_bins', 'ias_row_stride', 'ias_new_hist', 'ias_softmax_hist', 'ias_upsample_softmax_hist', 'ias_conf_hist', 'ias_threshold_scan',
    'ias_select', 'ias_meanprob_scan', 'ias_fused_window', 'UNSUPPORTED', 'cbst_sample_hist', 'cbst_quantile', 'copy_paste', 'hard_lut', 'st_loss_fwd', 'st_loss_bwd', 'st_loss_fused', 'st_loss_bwd_checked',
    'confusion_matrix', 'confusion_from_logits', 'iou_from_confusion', 'PngEncoder', 'resize_nearest_u8', 'softmax_flip_sum', 'probs_upsample_argmax', 'ce_general_fwd', 'ce_general_bwd', 'write_files',
    'Stager', 'FileWriter', 'WindowEmitter',
]


# ----------------------------------------------------------------------------- IAS
def ias_key_lo(num_classes):
    """Smallest fp16 key a softmax max-probability over C classes can take: fp16(1/C)."""
    k = lib().hiast_ias_key_lo(int(num_classes))
    check(min(k, 0), 'hiast_ias_key_lo')
    return k


def ias_num_bins(key_lo):
    return KEY_ONE - key_lo + 1


def ias_row_stride(key_lo):
    """Words per histogram row: the bins padded to a multiple of 4 (16-byte aligned rows)."""
    return lib().hiast_ias_hist_row_stride(int(key_lo))


def ias_new_hist(n_groups, num_classes, key_lo, device):
    """uint32 (as int32) [G, C, row_stride]; the first ias_num_bins(key_lo) entries of a row are its bins."""
    return torch.empty((n_groups, num_classes, ias_row_stride(key_lo)), dtype=torch.int32, device=device)


def _n_groups(n_images, group_size):
    return (n_images + group_size - 1) // group_size


def ias_softmax_hist(logits, group_size, key_lo=None, conf=None, label=None, hist=None, accumulate=False,
                     hist_mode=0):
    """Phase A.  logits f32 [N,C,H,W] -> (conf f32 [N,H,W], label u8 [N,H,W], hist i32 [G,C,nb])."""
    require_cuda(logits, torch.float32, 'logits')
    n, c, h, w = logits.shape
    if key_lo is None:
        key_lo = ias_key_lo(c)
    g = _n_groups(n, group_size)
    dev = logits.device
    conf = torch.empty((n, h, w), dtype=torch.float32, device=dev) if conf is None else require_cuda(conf, torch.float32, 'conf')
    label = torch.empty((n, h, w), dtype=torch.uint8, device=dev) if label is None else require_cuda(label, torch.uint8, 'label')
    if hist is None:
        hist = ias_new_hist(g, c, key_lo, dev)
        accumulate = False
    else:
        require_cuda(hist, torch.int32, 'hist')
        assert hist.numel() >= g * c * ias_row_stride(key_lo)
    if n == 0:
        return conf, label, hist
    check(lib().hiast_ias_softmax_hist(ptr(logits), n, c, h, w, int(group_size), int(key_lo), int(bool(accumulate)),
                                       int(hist_mode), ptr(conf), ptr(label), ptr(hist), stream_ptr(dev)),
          'hiast_ias_softmax_hist')
    return conf, label, hist


def ias_upsample_softmax_hist(logits_lr, out_hw, group_size, key_lo=None, conf=None, label=None, hist=None,
                              accumulate=False):
    """Phase A on LOW-RESOLUTION logits f32 [N,C,h,w]: bilinear (align_corners=True) up-sampling to out_hw fused in.
    Returns (conf f32 [N,H,W], label u8 [N,H,W], hist)."""
    require_cuda(logits_lr, torch.float32, 'logits_lr')
    n, c, h, w = logits_lr.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    if key_lo is None:
        key_lo = ias_key_lo(c)
    g = _n_groups(n, group_size)
    dev = logits_lr.device
    conf = torch.empty((n, H, W), dtype=torch.float32, device=dev) if conf is None else require_cuda(conf, torch.float32, 'conf')
    label = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if label is None else require_cuda(label, torch.uint8, 'label')
    if hist is None:
        hist = ias_new_hist(g, c, key_lo, dev)
        accumulate = False
    if n == 0:
        return conf, label, hist
    check(lib().hiast_ias_upsample_softmax_hist(ptr(logits_lr), n, c, h, w, H, W, int(group_size), int(key_lo),
                                                int(bool(accumulate)), ptr(conf), ptr(label), ptr(hist), stream_ptr(dev)),
          'hiast_ias_upsample_softmax_hist')
    return conf, label, hist


def ias_conf_hist(conf, label, num_classes, group_size, key_lo=0, hist=None, accumulate=False, label_u8_out=None):
    """Histogram from caller-provided conf f32 [N,H,W] and label (u8 or i64) [N,H,W].

    Returns (hist, label_u8): label_u8 is ``label`` itself when it already is uint8, else a uint8 copy
    (written into ``label_u8_out`` when given)."""
    require_cuda(conf, torch.float32, 'conf')
    require_cuda(label, (torch.uint8, torch.int64), 'label')
    assert conf.shape == label.shape
    n = conf.shape[0]
    hw = conf[0].numel() if n else 1
    g = _n_groups(n, group_size)
    dev = conf.device
    if hist is None:
        hist = ias_new_hist(g, num_classes, key_lo, dev)
        accumulate = False
    if label.dtype == torch.uint8:
        label_u8, out_ptr = label, None
        if label_u8_out is not None:
            label_u8_out.copy_(label)
            label_u8 = label_u8_out
    else:
        label_u8 = label_u8_out if label_u8_out is not None else torch.empty(label.shape, dtype=torch.uint8, device=dev)
        require_cuda(label_u8, torch.uint8, 'label_u8_out')
        out_ptr = ptr(label_u8)
    check(lib().hiast_ias_conf_hist(ptr(conf), ptr(label), label.element_size(), n, hw, int(num_classes),
                                    int(group_size), int(key_lo), int(bool(accumulate)), out_ptr, ptr(hist),
                                    stream_ptr(dev)), 'hiast_ias_conf_hist')
    return hist, label_u8


def ias_threshold_scan(hist, n_groups, num_classes, key_lo, alpha, beta, gamma, thr_state, thr_groups=None,
                       temp_groups=None, error_flag=None, token=None):
    """Phase B.  hist becomes prefix sums in place; thr_state f64[C] is updated in place.  ``token`` =
    (mailbox_in_ptr or None, in_seq, mailbox_out_ptr or None, out_seq): the multi-GPU hand-off over peer memory fused into
    the scan kernel (``hiast_ias_threshold_scan_ring``).

    Returns (thr_groups f64 [G,C], temp_groups f32 [G,C])."""
    require_cuda(hist, torch.int32, 'hist')
    require_cuda(thr_state, torch.float64, 'thr_state')
    dev = hist.device
    if thr_groups is None:
        thr_groups = torch.empty((n_groups, num_classes), dtype=torch.float64, device=dev)
    if temp_groups is None:
        temp_groups = torch.empty((n_groups, num_classes), dtype=torch.float32, device=dev)
    if token is None:
        check(lib().hiast_ias_threshold_scan(ptr(hist), int(n_groups), int(num_classes), int(key_lo), float(alpha),
                                             float(beta), float(gamma), ptr(thr_state), ptr(thr_groups), ptr(temp_groups),
                                             ptr(error_flag), stream_ptr(dev)), 'hiast_ias_threshold_scan')
    else:
        t_in, in_seq, t_out, out_seq = token
        check(lib().hiast_ias_threshold_scan_ring(ptr(hist), int(n_groups), int(num_classes), int(key_lo), float(alpha),
                                                  float(beta), float(gamma), ptr(thr_state), ptr(thr_groups), ptr(temp_groups),
                                                  ptr(error_flag), C.c_void_p(t_in) if t_in else None, int(in_seq),
                                                  C.c_void_p(t_out) if t_out else None, int(out_seq), stream_ptr(dev)),
              'hiast_ias_threshold_scan_ring')
    return thr_groups, temp_groups


def ias_select(conf, label, thr_groups, num_classes, group_size, plbl=None, counts=None, confsum=None):
    """Phase C.  Returns (plbl u8 [N,H,W], counts i64 [N,C], confsum i64(u64 bits) [G,C])."""
    require_cuda(conf, torch.float32, 'conf')
    require_cuda(label, torch.uint8, 'label')
    require_cuda(thr_groups, torch.float64, 'thr_groups')
    n = conf.shape[0]
    hw = conf[0].numel() if n else 1
    g = _n_groups(n, group_size)
    dev = conf.device
    if plbl is None:
        plbl = torch.empty(conf.shape, dtype=torch.uint8, device=dev)
    if counts is None:
        counts = torch.zeros((n, num_classes), dtype=torch.int64, device=dev)
    if confsum is None:
        confsum = torch.zeros((g, num_classes), dtype=torch.int64, device=dev)
    check(lib().hiast_ias_select(ptr(conf), ptr(label), ptr(thr_groups), n, hw, int(num_classes), int(group_size),
                                 ptr(plbl), ptr(counts), ptr(confsum), stream_ptr(dev)), 'hiast_ias_select')
    return plbl, counts, confsum


UNSUPPORTED = -2


def ias_fused_window(logits, group_size, key_lo, alpha, beta, gamma, conf, label, hist, thr_state, thr_groups, temp_groups,
                     plbl, counts, confsum, error_flag, workspace, keep_spill=False, groups_in_flight=0):
    """Phases A + B + C of a window in one persistent kernel (single GPU).  All buffers are caller-owned CUDA tensors
    (see include/hiast_b200.h).  Returns False when the shape is not covered by the fused kernel (nothing has been
    launched then: use the three-kernel path), True otherwise."""
    require_cuda(logits, torch.float32, 'logits')
    n, c, h, w = logits.shape
    require_cuda(conf, torch.float32, 'conf')
    require_cuda(label, torch.uint8, 'label')
    require_cuda(hist, torch.int32, 'hist')
    require_cuda(thr_state, torch.float64, 'thr_state')
    require_cuda(thr_groups, torch.float64, 'thr_groups')
    require_cuda(plbl, torch.uint8, 'plbl')
    require_cuda(counts, torch.int64, 'counts')
    require_cuda(confsum, torch.int64, 'confsum')
    require_cuda(error_flag, torch.int32, 'error_flag')
    require_cuda(workspace, torch.int32, 'workspace')
    g = _n_groups(n, group_size)
    assert conf.numel() >= n * h * w and label.numel() >= n * h * w and plbl.numel() >= n * h * w
    assert hist.numel() >= g * c * ias_row_stride(key_lo) and thr_groups.numel() >= g * c
    assert counts.numel() >= n * c and confsum.numel() >= g * c
    flags = (1 if keep_spill else 0) | ((int(groups_in_flight) & 0xf) << 4)
    status = lib().hiast_ias_fused_window(
        ptr(logits), n, c, h, w, int(group_size), int(key_lo), float(alpha), float(beta), float(gamma), ptr(conf), ptr(label),
        ptr(hist), ptr(thr_state), ptr(thr_groups), ptr(temp_groups), ptr(plbl), ptr(counts), ptr(confsum), ptr(error_flag),
        ptr(workspace), workspace.numel() * 4, flags, stream_ptr(logits.device))
    if status == UNSUPPORTED:
        return False
    check(status, 'hiast_ias_fused_window')
    return True


def ias_fused_workspace(n_images, group_size, device):
    nbytes = lib().hiast_ias_fused_workspace_bytes(int(n_images), int(group_size))
    return torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=device)


def ias_meanprob_scan(confsum, counts, group_size, num_classes, cp_gamma, mean_state):
    require_cuda(confsum, torch.int64, 'confsum')
    require_cuda(counts, torch.int64, 'counts')
    require_cuda(mean_state, torch.float64, 'mean_state')
    n = counts.shape[0]
    g = confsum.shape[0]
    check(lib().hiast_ias_meanprob_scan(ptr(confsum), ptr(counts), n, int(group_size), g, int(num_classes),
                                        float(cp_gamma), ptr(mean_state), stream_ptr(counts.device)),
          'hiast_ias_meanprob_scan')
    return mean_state


# ---------------------------------------------------------------------------- CBST
def cbst_sample_hist(conf, label, num_classes, group_size, sample_interval, key_lo, hist):
    """Accumulate the every-k-th-in-raster-order fp16 samples of one or more batches into hist i32 [C,row_stride]."""
    require_cuda(conf, torch.float32, 'conf')
    require_cuda(label, torch.uint8, 'label')
    require_cuda(hist, torch.int32, 'hist')
    n = conf.shape[0]
    hw = conf[0].numel() if n else 1
    dev = conf.device
    need = lib().hiast_cbst_workspace_bytes(n, hw, int(num_classes))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
    check(lib().hiast_cbst_sample_hist(ptr(conf), ptr(label), n, hw, int(num_classes), int(group_size), int(sample_interval),
                                       int(key_lo), ptr(hist), ptr(ws), ws.numel(), stream_ptr(dev)), 'hiast_cbst_sample_hist')
    return hist


def cbst_quantile(hist, num_classes, key_lo, q, error_flag=None):
    require_cuda(hist, torch.int32, 'hist')
    thr = torch.empty(num_classes, dtype=torch.float64, device=hist.device)
    check(lib().hiast_cbst_quantile(ptr(hist), int(num_classes), int(key_lo), float(q), ptr(thr), ptr(error_flag),
                                    stream_ptr(hist.device)), 'hiast_cbst_quantile')
    return thr


# ---------------------------------------------------------------------- copy-paste
def hard_lut(hard_classes):
    """256-bit set of class ids as 8 uint32 words (host)."""
    words = (C.c_uint32 * 8)()
    for c in hard_classes:
        c = int(c)
        if not 0 <= c < 256:
            raise ValueError('class id out of range: %r' % (c,))
        words[c >> 5] |= (1 << (c & 31))
    return words


def copy_paste(img, lbl, cp_mask, donor_img, donor_lbl, hard_classes, donor_index=None):
    """In place on img u8 [N,H,W,3], lbl u8 [N,H,W], cp_mask u8 [N,H,W]; donors indexed by donor_index (i32 [N])."""
    for t, name in ((img, 'img'), (lbl, 'lbl'), (cp_mask, 'cp_mask'), (donor_img, 'donor_img'), (donor_lbl, 'donor_lbl')):
        require_cuda(t, torch.uint8, name)
    n = lbl.shape[0]
    hw = lbl[0].numel() if n else 1
    assert img.numel() == n * hw * 3 and cp_mask.numel() == n * hw
    if donor_index is not None:
        require_cuda(donor_index, torch.int32, 'donor_index')
        assert donor_index.numel() == n
    else:
        assert donor_lbl.shape[0] >= n
    lut = hard_lut(hard_classes)
    check(lib().hiast_copy_paste(ptr(img), ptr(lbl), ptr(cp_mask), ptr(donor_img), ptr(donor_lbl), ptr(donor_index), n,
                                 hw, C.cast(lut, C.c_void_p), stream_ptr(lbl.device)), 'hiast_copy_paste')
    return img, lbl, cp_mask


# ---------------------------------------------------------------------------- loss
_ws_cache = {}


def _loss_workspace(b, c, hw, device):
    need = lib().hiast_st_loss_workspace_bytes(b, c, hw)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _plbl_arg(plbl):
    require_cuda(plbl, (torch.uint8, torch.int64), 'plbl')
    return plbl


def st_loss_fwd(z, t, plbl, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST):
    """Returns (sums f64[4], counts i64[3]) on the device; see include/hiast_b200.h."""
    require_cuda(z, torch.float32, 'logits')
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    if t is not None:
        require_cuda(t, torch.float32, 'target')
        assert t.shape == z.shape
    _plbl_arg(plbl)
    assert plbl.numel() == b * hw
    dev = z.device
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    counts = torch.empty(3, dtype=torch.int64, device=dev)
    ws = _loss_workspace(b, c, hw, dev)
    check(lib().hiast_st_loss_fwd(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region], int(terms),
                                  ptr(sums), ptr(counts), ptr(ws), ws.numel(), stream_ptr(dev)), 'hiast_st_loss_fwd')
    return sums, counts


def st_loss_bwd(z, t, plbl, scales, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST, grad=None):
    require_cuda(z, torch.float32, 'logits')
    require_cuda(scales, torch.float32, 'scales')
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    if grad is None:
        grad = torch.empty_like(z)
    check(lib().hiast_st_loss_bwd(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region], int(terms),
                                  ptr(scales), ptr(grad), stream_ptr(z.device)), 'hiast_st_loss_bwd')
    return grad


def st_loss_fused(z, t, plbl, grad_weights, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST, grad=None):
    """Forward + backward in one pass for ASSUMED upstream gradients ``grad_weights`` f32[4] (device).  Returns
    (sums f64[4], counts i64[3], scales_used f32[4], grad_z) or None when the configuration is not covered (nothing launched)."""
    require_cuda(z, torch.float32, 'logits')
    require_cuda(grad_weights, torch.float32, 'grad_weights')
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    if t is not None:
        require_cuda(t, torch.float32, 'target')
        assert t.shape == z.shape
    _plbl_arg(plbl)
    assert plbl.numel() == b * hw
    dev = z.device
    need = lib().hiast_st_loss_fused_workspace_bytes(b, c, hw)
    key = ('fused', dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = _ws_cache[key] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    counts = torch.empty(3, dtype=torch.int64, device=dev)
    used = torch.empty(4, dtype=torch.float32, device=dev)
    if grad is None:
        grad = torch.empty_like(z)
    status = lib().hiast_st_loss_fused(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region], int(terms),
                                       ptr(grad_weights), ptr(sums), ptr(counts), ptr(used), ptr(grad), ptr(ws), ws.numel(),
                                       stream_ptr(dev))
    if status == UNSUPPORTED:
        return None
    check(status, 'hiast_st_loss_fused')
    return sums, counts, used, grad


def st_loss_fused_terms(z, t, plbl, grad_weights, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST, grad=None,
                        term_weights=None):
    """``st_loss_fused`` with the four loss terms and their divisors computed on the device as well (hiast_st_loss_fused_terms).
    Returns (losses f32[4], divisors f64[4], scales_used f32[4], grad_z, sums f64[4], counts i64[3]) or None when the
    configuration is not covered.  The small outputs are slices of ONE allocation (the caller keeps them for backward).
    ``term_weights`` f32[4] (device): ``losses`` then holds the weighted terms w_k * loss_k."""
    require_cuda(z, torch.float32, 'logits')
    require_cuda(grad_weights, torch.float32, 'grad_weights')
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    if t is not None:
        require_cuda(t, torch.float32, 'target')
        assert t.shape == z.shape
    _plbl_arg(plbl)
    assert plbl.numel() == b * hw
    dev = z.device
    need = lib().hiast_st_loss_fused_workspace_bytes(b, c, hw)
    key = ('fused', dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = _ws_cache[key] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
    small = torch.empty(15, dtype=torch.float64, device=dev)          # sums 4 | counts 3 | divisors 4 | losses 4 f32 | used 4 f32
    sums, counts, divisors = small[0:4], small[4:7].view(torch.int64), small[7:11]
    f32 = small[11:15].view(torch.float32)
    losses, used = f32[0:4], f32[4:8]
    if grad is None:
        grad = torch.empty_like(z)
    status = lib().hiast_st_loss_fused_terms(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region], int(terms),
                                             ptr(grad_weights), ptr(sums), ptr(counts), ptr(used), ptr(grad), ptr(losses),
                                             ptr(divisors), ptr(term_weights), ptr(ws), ws.numel(), stream_ptr(dev))
    if status == UNSUPPORTED:
        return None
    check(status, 'hiast_st_loss_fused_terms')
    return losses, divisors, used, grad, sums, counts


def st_loss_bwd_checked_terms(z, t, plbl, gouts, divisors, scales_used, grad, region='ignored',
                              terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST, hint_weights=None, k0=0, upstream_out=None,
                              term_weights=None):
    """``st_loss_bwd_checked`` with the scales derived on the device from the upstream gradients ``gouts`` (four 0-d f32 CUDA
    tensors or None) and ``divisors`` f64[4]; optionally records gout[k0] / hint_weights[k0] in ``upstream_out``.  With
    ``term_weights`` the ``gouts`` are the gradients of the WEIGHTED terms (gout_k = gouts[k] * w_k)."""
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    gp = []
    for g in gouts:
        if g is not None:
            require_cuda(g, torch.float32, 'upstream gradient')
        gp.append(ptr(g))
    check(lib().hiast_st_loss_bwd_checked_terms(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region],
                                                int(terms), gp[0], gp[1], gp[2], gp[3], ptr(divisors), ptr(scales_used),
                                                ptr(grad), ptr(term_weights), ptr(hint_weights), int(k0), ptr(upstream_out),
                                                stream_ptr(z.device)),
          'hiast_st_loss_bwd_checked_terms')
    return grad


def st_loss_bwd_checked(z, t, plbl, scales, scales_used, grad, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST):
    """``grad`` (written by ``st_loss_fused`` for ``scales_used``) is left alone if ``scales`` are the same bits, else rewritten."""
    require_cuda(scales, torch.float32, 'scales')
    b, c = z.shape[0], z.shape[1]
    hw = z[0, 0].numel() if b else 1
    check(lib().hiast_st_loss_bwd_checked(ptr(z), ptr(t), ptr(plbl), plbl.element_size(), b, c, hw, REGION[region], int(terms),
                                          ptr(scales), ptr(scales_used), ptr(grad), stream_ptr(z.device)), 'hiast_st_loss_bwd_checked')
    return grad


# ----------------------------------------------------------------------- confusion
def confusion_matrix(pred, target, K, ignore_index=IGNORE, cm=None, mutate_pred=False):
    """cm i64 [K+1,K+1] accumulated; rows = target, cols = pred, index K = out of range."""
    require_cuda(pred, (torch.uint8, torch.int64), 'pred')
    require_cuda(target, (torch.uint8, torch.int64), 'target')
    assert pred.dtype == target.dtype and pred.numel() == target.numel()
    dev = pred.device
    if cm is None:
        cm = torch.zeros((K + 1, K + 1), dtype=torch.int64, device=dev)
    check(lib().hiast_confusion_matrix(ptr(pred), ptr(target), pred.element_size(), pred.numel(), int(K),
                                       int(ignore_index), ptr(pred) if mutate_pred else None, ptr(cm), stream_ptr(dev)),
          'hiast_confusion_matrix')
    return cm


def confusion_from_logits(logits, target, K, ignore_index=IGNORE, cm=None):
    require_cuda(logits, torch.float32, 'logits')
    require_cuda(target, (torch.uint8, torch.int64), 'target')
    b, c = logits.shape[0], logits.shape[1]
    hw = logits[0, 0].numel() if b else 1
    assert target.numel() == b * hw
    dev = logits.device
    if cm is None:
        cm = torch.zeros((K + 1, K + 1), dtype=torch.int64, device=dev)
    check(lib().hiast_confusion_from_logits(ptr(logits), ptr(target), target.element_size(), b, c, hw, int(K),
                                            int(ignore_index), ptr(cm), stream_ptr(dev)), 'hiast_confusion_from_logits')
    return cm


def iou_from_confusion(cm, K):
    require_cuda(cm, torch.int64, 'cm')
    inter = torch.empty(K, dtype=torch.float32, device=cm.device)
    union = torch.empty(K, dtype=torch.float32, device=cm.device)
    check(lib().hiast_iou_from_confusion(ptr(cm), int(K), ptr(inter), ptr(union), stream_ptr(cm.device)),
          'hiast_iou_from_confusion')
    return inter, union


# ----------------------------------------------------------------------------- PNG
class PngEncoder:
    """Device PNG writer for uint8 label maps [N,H,W] (pseudo_label_generator.py:43-46 `cv2.imwrite`).

    ``encode(labels)`` launches the three kernels on the current stream and returns ``(blob, offsets, host_offsets)``:
    device uint8 blob, device int64 [N+1] offsets and their pinned host copy (one sync); file i is
    ``blob[offsets[i]:offsets[i+1]]``.  ``encode_to_host(labels)`` also
    copies the used part of the blob to pinned host memory (one sync) and returns a list of ``memoryview``-able numpy
    slices.  The blob is sized for ``expect_ratio`` x compression and grown (to the worst case) if a batch does not fit."""

    def __init__(self, H, W, max_images, device='cuda', expect_ratio=4.0):
        self.H, self.W, self.max_images = int(H), int(W), int(max_images)
        self.device = torch.device(device)
        l = lib()
        self.max_file = l.hiast_png_max_bytes(self.H, self.W)
        if self.max_file == 0:
            raise _lib.HiastError('unsupported PNG size %dx%d' % (H, W))
        self.segments = l.hiast_png_segments(self.H, self.W)
        need = l.hiast_png_workspace_bytes(self.max_images, self.H, self.W)
        self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        cap = int(self.max_file * self.max_images / max(1.0, float(expect_ratio))) + 4096
        self._blob = torch.empty(cap, dtype=torch.uint8, device=self.device)
        self._offsets = torch.empty(self.max_images + 1, dtype=torch.int64, device=self.device)
        self._host = {}
        self._last_offsets = {}
        self._host_off = torch.empty(self.max_images + 1, dtype=torch.int64).pin_memory()

    def _launch(self, labels):
        n = labels.shape[0]
        check(lib().hiast_png_encode(ptr(labels), n, self.H, self.W, ptr(self._blob), self._blob.numel(), ptr(self._offsets),
                                     ptr(self._ws), self._ws.numel(), stream_ptr(self.device)), 'hiast_png_encode')

    def encode(self, labels):
        require_cuda(labels, torch.uint8, 'labels')
        if labels.dim() == 2:
            labels = labels.unsqueeze(0)
        n = labels.shape[0]
        if tuple(labels.shape[1:]) != (self.H, self.W) or n > self.max_images:
            raise _lib.HiastError('labels must be [<=%d, %d, %d], got %s' % (self.max_images, self.H, self.W, tuple(labels.shape)))
        self._launch(labels)
        self._host_off[:n + 1].copy_(self._offsets[:n + 1], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        total = int(self._host_off[n])
        if total > self._blob.numel():                  # did not fit: grow to the worst case and run again
            self._blob = torch.empty(self.max_file * self.max_images + 4096, dtype=torch.uint8, device=self.device)
            self._launch(labels)
        return self._blob[:total], self._offsets[:n + 1], self._host_off[:n + 1]

    def encode_to_host(self, labels, slot=0):
        """List of numpy uint8 arrays, one PNG file each: views of the pinned buffer `slot`, valid until the next call
        with the same slot (two slots let a writer thread pool drain one window while the next is encoded)."""
        blob, _, off = self.encode(labels)
        total = blob.numel()
        host = self._host.get(slot)
        if host is None or host.numel() < total:
            host = self._host[slot] = torch.empty(max(total + total // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        host[:total].copy_(blob, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        h = host.numpy()
        o = off.tolist()
        self._last_offsets[slot] = o
        return [h[o[i]:o[i + 1]] for i in range(len(o) - 1)]


def resize_nearest_u8(labels, size):
    """``cv2.resize(lbl, (W, H), interpolation=cv2.INTER_NEAREST)`` (base_dataset.py:176) for uint8 label maps
    [N,Hs,Ws] (or [Hs,Ws]) on the device -> [N,H,W]."""
    require_cuda(labels, torch.uint8, 'labels')
    single = labels.dim() == 2
    src = labels.unsqueeze(0) if single else labels
    n, hs, ws = src.shape
    hd, wd = int(size[0]), int(size[1])
    dst = torch.empty((n, hd, wd), dtype=torch.uint8, device=labels.device)
    ifx = 1.0 / (float(wd) / float(ws))             # OpenCV: inv_scale = dsize / ssize; scale = 1. / inv_scale
    ify = 1.0 / (float(hd) / float(hs))
    check(lib().hiast_resize_nearest_u8(ptr(src), n, hs, ws, ptr(dst), hd, wd, ifx, ify, stream_ptr(labels.device)),
          'hiast_resize_nearest_u8')
    return dst[0] if single else dst


# ----------------------------------------------------------------------- validator
def softmax_flip_sum(logits, logits_of_flipped=None, out=None):
    """softmax(logits, 1) [+ flip_x(softmax(logits_of_flipped, 1))] for f32 [B,C,h,w] (validator.py:37,48-50)."""
    require_cuda(logits, torch.float32, 'logits')
    if logits_of_flipped is not None:
        require_cuda(logits_of_flipped, torch.float32, 'logits_of_flipped')
        assert logits_of_flipped.shape == logits.shape
    b, c, h, w = logits.shape
    if out is None:
        out = torch.empty_like(logits)
    check(lib().hiast_softmax_flip_sum(ptr(logits), ptr(logits_of_flipped), b, c, h, w, ptr(out), stream_ptr(logits.device)),
          'hiast_softmax_flip_sum')
    return out


def probs_upsample_argmax(probs_list, size):
    """uint8 [B,H,W] = argmax_c sum_s interpolate(probs_s, size, bilinear, align_corners=True) (validator.py:52-55,93)."""
    n = len(probs_list)
    for p in probs_list:
        require_cuda(p, torch.float32, 'probs')
    b, c = probs_list[0].shape[:2]
    assert all(p.shape[:2] == (b, c) for p in probs_list)
    H, W = int(size[0]), int(size[1])
    label = torch.empty((b, H, W), dtype=torch.uint8, device=probs_list[0].device)
    ptrs = (C.c_void_p * n)(*[p.data_ptr() for p in probs_list])
    hs = (C.c_int * n)(*[p.shape[2] for p in probs_list])
    ws = (C.c_int * n)(*[p.shape[3] for p in probs_list])
    check(lib().hiast_probs_upsample_argmax(C.cast(ptrs, C.c_void_p), C.cast(hs, C.c_void_p), C.cast(ws, C.c_void_p), n, b, c,
                                            H, W, ptr(label), stream_ptr(label.device)), 'hiast_probs_upsample_argmax')
    return label


# ------------------------------------------------------ CE with class weights / refer_labels
def _ce_general_args(z, labels, weights, refer_labels, region):
    require_cuda(z, torch.float32, 'logits')
    require_cuda(labels, (torch.uint8, torch.int64), 'labels')
    if weights is not None:
        require_cuda(weights, torch.float32, 'weights')
    if refer_labels is not None:
        require_cuda(refer_labels, (torch.uint8, torch.int64), 'refer_labels')
    b, c = z.shape[:2]
    hw = z[0, 0].numel() if b else 1
    return (ptr(z), ptr(labels), labels.element_size(), ptr(weights), ptr(refer_labels),
            refer_labels.element_size() if refer_labels is not None else 0, REGION.get(region, 1)), b, c, hw


def ce_general_fwd(z, labels, weights=None, refer_labels=None, region='confident', ignore_index=IGNORE):
    """(sums f64[2], count i64[1]) of hiast_ce_general_fwd (losses.py:32-36 with weights / refer_labels)."""
    head, b, c, hw = _ce_general_args(z, labels, weights, refer_labels, region)
    sums = torch.empty(2, dtype=torch.float64, device=z.device)
    count = torch.empty(1, dtype=torch.int64, device=z.device)
    ws = torch.empty(max(lib().hiast_ce_general_workspace_bytes(hw), 8), dtype=torch.uint8, device=z.device)
    check(lib().hiast_ce_general_fwd(*head, int(ignore_index), b, c, hw, ptr(sums), ptr(count), ptr(ws), ws.numel(),
                                     stream_ptr(z.device)), 'hiast_ce_general_fwd')
    return sums, count


def ce_general_bwd(z, labels, weights, refer_labels, region, ignore_index, scale):
    head, b, c, hw = _ce_general_args(z, labels, weights, refer_labels, region)
    require_cuda(scale, torch.float32, 'scale')
    grad = torch.empty_like(z)
    check(lib().hiast_ce_general_bwd(*head, int(ignore_index), b, c, hw, ptr(scale), ptr(grad), stream_ptr(z.device)),
          'hiast_ce_general_bwd')
    return grad


def write_files(paths, blob_host, offsets, n_threads=8):
    """Write file i = blob_host[offsets[i]:offsets[i+1]] to paths[i] with native POSIX writer threads (one foreign call,
    the interpreter lock is released for its duration).  blob_host: contiguous uint8 numpy array."""
    n = len(paths)
    assert len(offsets) == n + 1 and blob_host.dtype == np.uint8 and blob_host.flags['C_CONTIGUOUS']
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    off = (C.c_int64 * (n + 1))(*[int(o) for o in offsets])
    err = C.c_int(0)
    status = lib().hiast_write_files(C.cast(arr, C.c_void_p), C.c_void_p(blob_host.ctypes.data), C.cast(off, C.c_void_p), n,
                                     int(n_threads), C.cast(C.pointer(err), C.c_void_p))
    if status != 0:
        raise OSError(err.value, 'hiast_write_files: %s' % os.strerror(err.value) if err.value else 'hiast_write_files failed')


# ------------------------------------------------------------ host pipeline (csrc/host_pipeline.cu)
class Stager:
    """Host-to-device staging ring (``hiast_stager_*``): ``n_slots`` device slots of ``slot_bytes`` carved from one uint8 buffer,
    filled by ``cudaMemcpyAsync`` on a copy stream.  ``push`` returns a device view (same dtype / shape as the host tensor) that
    the consumer stream may use at once -- the ordering events are queued, the host never waits.  ``release`` marks slots as
    consumed by everything queued on the consumer stream so far; a slot must be released before it is pushed again."""

    def __init__(self, n_slots, slot_bytes, device, copy_stream):
        self.n_slots, self.device, self.copy_stream = int(n_slots), torch.device(device), copy_stream
        # slots are packed back to back (4-byte granularity): consecutive stride-8 batches of a window form ONE contiguous
        # [n, C, h, w] tensor in the ring, so their single phase-A launch needs no torch.cat (a 161 MB copy per window)
        self.slot_bytes = (int(slot_bytes) + 3) // 4 * 4
        self.ring = torch.empty(self.n_slots * self.slot_bytes, dtype=torch.uint8, device=self.device)
        self._busy = [False] * self.n_slots
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().hiast_stager_create(self.n_slots, C.byref(h)), 'hiast_stager_create')
        self._h = h
        self._push, self._release = lib().hiast_stager_push, lib().hiast_stager_release
        self._base = self.ring.data_ptr()
        self._cs = C.c_void_p(copy_stream.cuda_stream)

    def busy(self, slot):
        return self._busy[slot]

    def push(self, slot, host_tensor, consumer_stream_ptr):
        if self._busy[slot]:
            raise _lib.HiastError('staging slot %d pushed again before it was released' % slot)
        t = host_tensor if host_tensor.is_contiguous() else host_tensor.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes > self.slot_bytes:
            raise _lib.HiastError('batch of %d bytes does not fit a %d-byte staging slot' % (nbytes, self.slot_bytes))
        off = slot * self.slot_bytes
        check(self._push(self._h, slot, C.c_void_p(self._base + off), C.c_void_p(t.data_ptr()), nbytes, self._cs,
                         consumer_stream_ptr), 'hiast_stager_push')
        self._busy[slot] = True
        return self.ring[off:off + nbytes].view(t.dtype).view(t.shape)

    def release(self, first_slot, n_slots, consumer_stream_ptr):
        check(self._release(self._h, first_slot, n_slots, consumer_stream_ptr), 'hiast_stager_release')
        for k in range(first_slot, first_slot + n_slots):
            self._busy[k] = False

    def view_of(self, first_ptr, nbytes, dtype, shape):
        """A tensor over ring memory [first_ptr, first_ptr + nbytes) (consecutive slots of one window), else None."""
        off = first_ptr - self._base
        if off < 0 or off + nbytes > self.ring.numel():
            return None
        return self.ring[off:off + nbytes].view(dtype).view(shape)

    def close(self):
        if self._h is not None:
            lib().hiast_stager_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FileWriter:
    """Persistent native writer pool (``hiast_writer_*``).  ``submit`` returns a ticket immediately; the files are on disk
    when ``wait(ticket)`` returns (foreign call, interpreter lock released)."""

    def __init__(self, n_threads, device):
        self.device = torch.device(device)
        h = C.c_void_p()
        check(lib().hiast_writer_create(max(1, int(n_threads)), C.byref(h)), 'hiast_writer_create')
        self._h = h
        self.n_threads = max(1, int(n_threads))

    def submit(self, paths, blob_host, offsets_host, bytes_copied, blob_dev, stream_ptr_):
        n = len(paths)
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        t = lib().hiast_writer_submit(self._h, C.cast(arr, C.c_void_p), n, C.c_void_p(blob_host.data_ptr()), blob_host.numel(),
                                      C.c_void_p(offsets_host.data_ptr()), int(bytes_copied), C.c_void_p(blob_dev.data_ptr()),
                                      stream_ptr_)
        if t <= 0:
            check(int(t) if t < 0 else -1, 'hiast_writer_submit')
        return int(t)

    def wait(self, ticket):
        err = C.c_int(0)
        status = lib().hiast_writer_wait(self._h, int(ticket), C.byref(err))
        if status == -5:
            raise OSError(err.value, 'pseudo-label file writer: %s' % os.strerror(err.value))
        check(status, 'hiast_writer_wait')

    def close(self):
        if self._h is not None:
            lib().hiast_writer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WindowEmitter:
    """Per-window outputs through ONE foreign call (``hiast_ias_emit_window``): phase C, optional mean-prob EMA, the PNG
    encoder (``png=True``) or the raw uint8 label maps (``png=False``), and every device-to-host copy, into the pinned
    buffers of one of ``n_slots`` emit slots.  The caller orders reuse of a slot (writer ticket or event)."""

    def __init__(self, engine, window_images, n_slots=3, png=True):
        e = self.engine = engine
        self.in_use = False
        self.window, self.n_slots, self.png = int(window_images), int(n_slots), bool(png)
        self.device = e.device
        n, g, c, hw = self.window, (self.window + e.B - 1) // e.B, e.C, e.H * e.W
        l = lib()
        self.slots = []
        self.max_file = l.hiast_png_max_bytes(e.H, e.W) if png else 0
        if png and self.max_file == 0:
            raise _lib.HiastError('unsupported PNG size %dx%d' % (e.H, e.W))
        ws_bytes = l.hiast_png_workspace_bytes(n, e.H, e.W) if png else 0
        self._png_ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=self.device) if png else None
        cap = self.max_file * n + 4096
        self.predicted = cap // 4 + 4096                    # bytes of a window's files; refined from every finished window
        for _ in range(self.n_slots):
            s = dict(counts_host=torch.empty((n, c), dtype=torch.int64).pin_memory(),
                     confsum_host=torch.empty((g, c), dtype=torch.int64).pin_memory(),
                     thr_host=torch.empty((g, c), dtype=torch.float64).pin_memory())
            if png:
                s.update(blob_dev=torch.empty(cap, dtype=torch.uint8, device=self.device),
                         offsets_dev=torch.empty(n + 1, dtype=torch.int64, device=self.device),
                         blob_host=torch.empty(cap, dtype=torch.uint8).pin_memory(),
                         offsets_host=torch.zeros(n + 1, dtype=torch.int64).pin_memory())
            else:
                s.update(plbl_host=torch.empty((n, e.H, e.W), dtype=torch.uint8).pin_memory())
            a = _lib.WindowEmit()
            a.H, a.W, a.C, a.group_size, a.cp_gamma = e.H, e.W, e.C, e.B, e.cp_gamma
            a.counts_host, a.confsum_host, a.thr_groups_host = (s['counts_host'].data_ptr(), s['confsum_host'].data_ptr(),
                                                                s['thr_host'].data_ptr())
            if png:
                a.blob_dev, a.offsets_dev, a.png_ws = s['blob_dev'].data_ptr(), s['offsets_dev'].data_ptr(), self._png_ws.data_ptr()
                a.blob_host, a.offsets_host = s['blob_host'].data_ptr(), s['offsets_host'].data_ptr()
                a.blob_capacity, a.png_ws_bytes = cap, self._png_ws.numel()
            else:
                a.plbl_host = s['plbl_host'].data_ptr()
            s['args'] = a
            self.slots.append(s)
        self._emit = l.hiast_ias_emit_window

    _CACHE = {}

    @classmethod
    def get(cls, engine, window_images, n_slots=3, png=True):
        """An emitter for this engine geometry, reused across generator runs of one process (its pinned and device buffers
        -- several hundred MB at full resolution -- take ~10 ms to allocate and pin)."""
        key = (str(engine.device), engine.C, engine.H, engine.W, engine.B, float(engine.cp_gamma), int(window_images), int(n_slots),
               bool(png))
        em = cls._CACHE.get(key)
        if em is None or em.in_use:                      # in use: another live pipeline owns its buffers
            em = cls(engine, window_images, n_slots, png)
            if key not in cls._CACHE:
                if len(cls._CACHE) >= 4:
                    cls._CACHE.clear()
                cls._CACHE[key] = em
        em.engine = engine
        em.in_use = True
        return em

    def done(self):
        """The pipeline that took this emitter from ``get`` has completed every window."""
        self.in_use = False

    def emit(self, slot, first_image, n_images, with_mean_prob=False, stream=None, copy_stream=None):
        """Queue the outputs of engine images [first_image, first_image + n) into emit slot ``slot`` on the current stream."""
        e, s = self.engine, self.slots[slot]
        a = s['args']
        g0 = first_image // e.B
        a.conf = e.conf.data_ptr() + first_image * e.H * e.W * 4
        a.label = e.label.data_ptr() + first_image * e.H * e.W
        a.plbl = e.plbl.data_ptr() + first_image * e.H * e.W
        a.thr_groups = e.thr_groups.data_ptr() + g0 * e.C * 8
        a.counts = e.counts.data_ptr() + first_image * e.C * 8
        a.confsum = e.confsum.data_ptr() + g0 * e.C * 8
        a.mean_state = e.mean_state.data_ptr() if with_mean_prob else None
        a.n_images = n_images
        copied = 0
        if self.png:
            copied = a.blob_copy_bytes = min(self.predicted, s['blob_dev'].numel())
        check(self._emit(C.byref(a), stream if stream is not None else stream_ptr(self.device), copy_stream),
              'hiast_ias_emit_window')
        return copied

    def learn(self, total_bytes):
        """Feed back the size of a finished window's files (the next windows copy 1.25 x that with their results)."""
        self.predicted = int(total_bytes) + int(total_bytes) // 4 + 65536
