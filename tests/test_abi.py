"""The C-ABI library loads without a GPU and exports every symbol include/hiast_b200.h (the drop-in boundary) and
include/hiast_b200_dev.h (test hooks / A-B toggles) declare."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'hiast_b200.h')
DEV_HEADER = os.path.join(ROOT, 'include', 'hiast_b200_dev.h')


@pytest.fixture(scope='module')
def built_lib():
    from hiast_b200 import build
    path = build.build()
    assert os.path.exists(path)
    return path


def declared_functions(headers=(HEADER, DEV_HEADER)):
    names = []
    for h in headers:
        text = open(h).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names += re.findall(r'\b(hiast_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_public_header_holds_no_debug_or_test_symbols():
    public = declared_functions((HEADER,))
    assert not [n for n in public if 'debug' in n or 'testhook' in n or 'selftest' in n or n.startswith('hiast_dev_')]
    dev = set(declared_functions((DEV_HEADER,))) - set(public)
    assert all('debug' in n or 'testhook' in n or 'selftest' in n or n.startswith('hiast_dev_') for n in dev), dev


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ['hiast_ias_softmax_hist', 'hiast_ias_threshold_scan', 'hiast_ias_select', 'hiast_ias_meanprob_scan',
                 'hiast_copy_paste', 'hiast_st_loss_fwd', 'hiast_st_loss_bwd', 'hiast_confusion_matrix']:
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    missing = [n for n in declared_functions() if not hasattr(handle, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(built_lib):
    from hiast_b200 import _lib
    assert sorted(_lib.exported_symbols()) == declared_functions()
    l = _lib.lib()
    assert l.hiast_version() >= 1
    assert l.hiast_status_string(0) == b'ok'
    assert l.hiast_status_string(-1) == b'invalid argument'


def test_argument_validation_without_gpu(built_lib):
    """Invalid-argument paths return before touching CUDA."""
    from hiast_b200 import _lib
    l = _lib.lib()
    assert l.hiast_ias_key_lo(19) == 0x2ABD
    assert l.hiast_ias_key_lo(0) == -1
    assert l.hiast_ias_hist_row_stride(0x2ABD) == 4420 and l.hiast_ias_hist_row_stride(0) == 15364
    assert l.hiast_ias_hist_bytes(3, 19, 0x2ABD) == 3 * 19 * 4420 * 4
    assert l.hiast_ias_softmax_hist(None, 1, 19, 4, 4, 2, 0, 0, 0, None, None, None, None) == -1
    assert l.hiast_st_loss_fwd(None, None, None, 8, 1, 19, 4, 0, 15, None, None, None, 0, None) == -1
    assert l.hiast_confusion_matrix(None, None, 8, 4, 19, 255, None, None, None) == -1
    assert l.hiast_copy_paste(None, None, None, None, None, None, 1, 4, None, None) == -1


def test_product_path_has_no_cpu_fallback(built_lib):
    import torch
    from hiast_b200 import _lib, ops
    with pytest.raises(_lib.HiastError):
        ops.ias_softmax_hist(torch.zeros(1, 19, 4, 4), 2)
    with pytest.raises(_lib.HiastError):
        ops.confusion_matrix(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 19)
    with pytest.raises(_lib.HiastError):
        ops.ias_upsample_softmax_hist(torch.zeros(1, 19, 3, 3), (8, 8), 2)
    with pytest.raises(_lib.HiastError):                       # round-1 widening: no host path either
        ops.resize_nearest_u8(torch.zeros(2, 4, 4, dtype=torch.uint8), (8, 8))
    with pytest.raises(_lib.HiastError):
        ops.softmax_flip_sum(torch.zeros(1, 19, 4, 4))
    with pytest.raises(_lib.HiastError):
        ops.probs_upsample_argmax([torch.zeros(1, 19, 4, 4)], (8, 8))
    with pytest.raises(_lib.HiastError):
        ops.ce_general_fwd(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
    from hiast_b200.losses import ce
    with pytest.raises(_lib.HiastError):
        ce(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64), weights=torch.ones(19))
    from hiast_b200.validator import Validator
    from types import SimpleNamespace
    cfg = SimpleNamespace(dataset=SimpleNamespace(num_classes=19, source=SimpleNamespace(type='GTA5')),
                          validate=SimpleNamespace(resize_sizes=[[4, 4]], is_flip=False, color_mask_dir_path=None))
    with pytest.raises(_lib.HiastError):
        Validator(cfg, model=lambda x: {'logits': torch.zeros(1, 19, 4, 4)}, loader=[], device='cpu').predict(torch.zeros(1, 3, 4, 4))
    from hiast_b200.ema import update_ema_model
    with pytest.raises(_lib.HiastError):                       # CPU models: no host path for the EMA update either
        update_ema_model(torch.nn.Linear(3, 2), torch.nn.Linear(3, 2), 0.99)
    l = _lib.lib()
    assert l.hiast_ema_update(None, None, None, 1, 1024, 0.9, 0.1, None) == -1
    assert l.hiast_png_encode(None, 1, 4, 4, None, 0, None, None, 0, None) == -1
    assert l.hiast_resize_nearest_u8(None, 1, 4, 4, None, 8, 8, 1.0, 1.0, None) == -1
    assert l.hiast_softmax_flip_sum(None, None, 1, 19, 4, 4, None, None) == -1
    assert l.hiast_probs_upsample_argmax(None, None, None, 1, 1, 19, 4, 4, None, None) == -1
    assert l.hiast_ce_general_fwd(None, None, 8, None, None, 0, 1, 255, 1, 19, 16, None, None, None, 0, None) == -1
    assert l.hiast_write_files(None, None, None, 1, 1, None) == -1
    assert l.hiast_ias_fused_window(None, 1, 19, 4, 4, 2, 0, 0.5, 0.9, 8.0, None, None, None, None, None, None, None, None, None,
                                    None, None, 0, 0, None) == -1


def test_native_file_writer_is_a_host_function(built_lib, tmp_path):
    """hiast_write_files (the writer behind the device PNG path) needs no GPU: contents, empty files, error reporting."""
    import numpy as np
    from hiast_b200 import ops
    blob = np.frombuffer(os.urandom(20000), dtype=np.uint8).copy()
    offs = [0, 7, 7, 9000, 20000]
    paths = [str(tmp_path / ('f%d_pseudo_label.png' % i)) for i in range(4)]
    ops.write_files(paths, blob, offs, n_threads=3)
    for i, p in enumerate(paths):
        assert open(p, 'rb').read() == blob[offs[i]:offs[i + 1]].tobytes()
    ops.write_files(paths[:1], blob, [5, 6], n_threads=8)                 # truncates an existing file
    assert open(paths[0], 'rb').read() == blob[5:6].tobytes()
    with pytest.raises(OSError) as e:
        ops.write_files([str(tmp_path / 'missing_dir' / 'x.png')], blob, [0, 4])
    assert e.value.errno == 2


def test_lean_loss_entry_points_validate_and_have_no_cpu_path(built_lib):
    """hiast_st_loss_fused_terms / hiast_st_loss_bwd_checked_terms: invalid arguments return before touching CUDA; the Python
    wrappers refuse CPU tensors; FusedTermsLean is only taken for configurations the one-pass kernel covers."""
    import torch
    from hiast_b200 import _lib, ops
    from hiast_b200 import losses as L
    l = _lib.lib()
    assert l.hiast_st_loss_fused_terms(None, None, None, 8, 1, 19, 4, 0, 15, None, None, None, None, None, None, None, None, None, 0,
                                       None) == -1
    assert l.hiast_st_loss_bwd_checked_terms(None, None, None, 8, 1, 19, 4, 0, 15, None, None, None, None, None, None, None, None,
                                             None, 0, None, None) == -1
    z = torch.zeros(1, 19, 4, 4)
    y = torch.zeros(1, 4, 4, dtype=torch.int64)
    gw = torch.ones(4)
    with pytest.raises(_lib.HiastError):
        ops.st_loss_fused_terms(z, z, y, gw)
    with pytest.raises(_lib.HiastError):
        ops.st_loss_bwd_checked_terms(z, z, y, [torch.ones(())] * 4, torch.ones(4, dtype=torch.float64), gw, z)
    # which configurations take the lean path
    hint = object()
    zg = torch.zeros(2, 19, 4, 6, requires_grad=True)
    assert L.lean_ok(zg, y, 15, False, hint)
    assert not L.lean_ok(zg, y, 15, False, None)                          # no expectation of the upstream gradient
    assert not L.lean_ok(zg, y, 15, True, hint)                           # SoftCE without refer_labels: divisor = numel
    assert not L.lean_ok(zg, y, 15 | _lib.CST_KLDIV, False, hint)         # other consistency kinds
    assert not L.lean_ok(torch.zeros(2, 19, 4, 6), y, 15, False, hint)    # no gradient wanted
    assert not L.lean_ok(torch.zeros(2, 7, 4, 6, requires_grad=True), y, 15, False, hint)    # C outside {16, 19}
    assert not L.lean_ok(torch.zeros(2, 19, 3, 5, requires_grad=True), y, 15, False, hint)   # odd HW
    with torch.no_grad():
        assert not L.lean_ok(zg, y, 15, False, hint)
