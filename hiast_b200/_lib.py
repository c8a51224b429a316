"""ctypes binding of libhiast_b200.so (the C ABI declared in include/hiast_b200.h).

There is no CPU fallback: if the library has not been built, or a call fails, this raises.
torch is used only as the owner of device memory and streams (``data_ptr()``, current stream).
"""

from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhiast_b200.so')

OK = 0
REGION = {'ignored': 0, 'confident': 1, 'all': 2}
TERM_CE, TERM_KLD, TERM_ENT, TERM_CST = 1, 2, 4, 8
CST_SOFTCE, CST_KLDIV, CST_MSE, CST_SOFTCE_LOGITS = 0, 16, 32, 48
KEY_ONE = 0x3C00
IGNORE = 255

_vp, _i, _i64, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double



class WindowEmit(C.Structure):
    """HiastWindowEmit of include/hiast_b200.h (argument block of hiast_ias_emit_window)."""
    _fields_ = [(n, _vp) for n in ('conf', 'label', 'thr_groups', 'plbl', 'counts', 'confsum', 'mean_state', 'blob_dev',
                                   'offsets_dev', 'png_ws', 'blob_host', 'offsets_host', 'plbl_host', 'counts_host',
                                   'confsum_host', 'thr_groups_host')] + \
               [('blob_capacity', _sz), ('png_ws_bytes', _sz), ('blob_copy_bytes', _sz), ('cp_gamma', _d)] + \
               [(n, C.c_int32) for n in ('n_images', 'H', 'W', 'C', 'group_size', 'reserved')]


_SIGNATURES = {
    'hiast_version': (_i, []),
    'hiast_status_string': (C.c_char_p, [_i]),
    'hiast_last_cuda_error': (_i, []),
    'hiast_device_sm_count': (_i, []),
    'hiast_ias_key_lo': (_i, [_i]),
    'hiast_ias_hist_row_stride': (_i, [_i]),
    'hiast_ias_hist_bytes': (_sz, [_i, _i, _i]),
    'hiast_ias_softmax_hist': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'hiast_ias_upsample_softmax_hist': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'hiast_ias_conf_hist': (_i, [_vp, _vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _vp, _vp]),
    'hiast_ias_threshold_scan': (_i, [_vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp]),
    'hiast_ias_threshold_scan_ring': (_i, [_vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp]),
    'hiast_ring_mailbox_bytes': (_sz, []),
    'hiast_ring_create': (_i, [C.POINTER(_vp), _vp]),
    'hiast_ring_open': (_i, [_vp, C.POINTER(_vp)]),
    'hiast_ring_close': (_i, [_vp]),
    'hiast_ring_destroy': (_i, [_vp]),
    'hiast_ias_select': (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    'hiast_ias_meanprob_scan': (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp]),
    'hiast_ias_fused_workspace_bytes': (_sz, [_i, _i]),
    'hiast_ias_fused_window': (_i, [_vp, _i, _i, _i, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _sz, _i, _vp]),
    'hiast_cbst_workspace_bytes': (_sz, [_i, _i64, _i]),
    'hiast_cbst_sample_hist': (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    'hiast_cbst_quantile': (_i, [_vp, _i, _i, _d, _vp, _vp, _vp]),
    'hiast_copy_paste': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    'hiast_st_loss_workspace_bytes': (_sz, [_i, _i, _i64]),
    'hiast_st_loss_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    'hiast_st_loss_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp]),
    'hiast_st_loss_fused_workspace_bytes': (_sz, [_i, _i, _i64]),
    'hiast_st_loss_fused': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'hiast_st_loss_bwd_checked': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    'hiast_st_loss_fused_terms': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                       _vp]),
    'hiast_st_loss_bwd_checked_terms': (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                             _i, _vp, _vp]),
    'hiast_confusion_matrix': (_i, [_vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp]),
    'hiast_confusion_from_logits': (_i, [_vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp]),
    'hiast_iou_from_confusion': (_i, [_vp, _i, _vp, _vp, _vp]),
    'hiast_ema_update': (_i, [_vp, _vp, _vp, _i, _i, C.c_float, C.c_float, _vp]),
    'hiast_multi_copy': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'hiast_png_workspace_bytes': (_sz, [_i, _i, _i]),
    'hiast_png_max_bytes': (_sz, [_i, _i]),
    'hiast_png_segments': (_i, [_i, _i]),
    'hiast_write_files': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'hiast_stager_create': (_i, [_i, C.POINTER(_vp)]),
    'hiast_stager_destroy': (_i, [_vp]),
    'hiast_stager_push': (_i, [_vp, _i, _vp, _vp, _sz, _vp, _vp]),
    'hiast_stager_release': (_i, [_vp, _i, _i, _vp]),
    'hiast_ias_emit_window': (_i, [C.POINTER(WindowEmit), _vp, _vp]),
    'hiast_writer_create': (_i, [_i, C.POINTER(_vp)]),
    'hiast_writer_destroy': (_i, [_vp]),
    'hiast_writer_submit': (_i64, [_vp, _vp, _i, _vp, _sz, _vp, _sz, _vp, _vp]),
    'hiast_writer_wait': (_i, [_vp, _i64, C.POINTER(_i)]),
    'hiast_dev_variants': (_i, []),
    'hiast_png_encode': (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, _vp, _sz, _vp]),
    'hiast_resize_nearest_u8': (_i, [_vp, _i, _i, _i, _vp, _i, _i, _d, _d, _vp]),
    'hiast_softmax_flip_sum': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    'hiast_probs_upsample_argmax': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'hiast_debug_validate_direct': (_i, [_i]),
    'hiast_ce_general_workspace_bytes': (_sz, [_i64]),
    'hiast_ce_general_fwd': (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    'hiast_ce_general_bwd': (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _vp]),
    'hiast_debug_png_variant': (_i, [_i]),
    'hiast_debug_set_fused_trace': (_i, [_vp]),
    'hiast_debug_loss_scalar': (_i, [_i]),
    'hiast_debug_upsample_v1': (_i, [_i]),
    'hiast_selftest_packed_expf': (_i, [_vp, _vp]),
    'hiast_testhook_powi': (_d, [_d, _i]),
    'hiast_testhook_threshold_step': (_d, [_vp, _i, _d, _d, _d, _d, _vp, _vp]),
}

_lib = None


class HiastError(RuntimeError):
    status = None          # the HIAST_ERR_* code when the error came from a C-ABI call


def lib():
    """The loaded library; raises if it has not been built (python -m hiast_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HiastError('%s is missing: build it with `python -m hiast_b200.build` '
                             '(there is no CPU fallback)' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(status, what):
    if status != OK:
        l = lib()
        msg = l.hiast_status_string(status).decode()
        if status == -3:
            msg += ' [cudaError %d]' % l.hiast_last_cuda_error()
        err = HiastError('%s failed: %s' % (what, msg))
        err.status = status
        raise err


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HiastError('%s must be a CUDA tensor (there is no CPU path)' % name)
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise HiastError('%s must have dtype %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise HiastError('%s must be contiguous' % name)
    return t
