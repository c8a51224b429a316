#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402
from hiast_b200.ema import update_ema_model  # noqa: E402
from hiast_b200.ias_engine import IASEngine  # noqa: E402

g = torch.Generator().manual_seed(3)
C, H, W, B, N = 19, 64, 128, 2, 5
logits = (torch.randn(N, C, H, W, generator=g) * 3).cuda()
logits[1, 4] += 30.0                                         # saturated pixels: top-key counters
from hiast_b200 import _lib  # noqa: E402
DEV = bool(_lib.lib().hiast_dev_variants())                  # development build: the dropped variants are in the library too
for fused in ((False, True) if DEV else (False,)):
    eng = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 6, fused=fused)
    eng.process(logits)
    assert eng.check_errors()
for mode in ((1, 6, 16, 36, 56, 80, 81, 83) if DEV else (1, 83)):
    ops.ias_softmax_hist(logits, B, hist_mode=mode)
lr = (torch.randn(3, C, 17, 33, generator=g) * 4).cuda()
ops.ias_upsample_softmax_hist(lr, (128, 256), B)
z = (torch.randn(2, C, 32, 64, generator=g) * 3).cuda()
t = torch.softmax(torch.randn(2, C, 32, 64, generator=g) * 3, dim=1).cuda()
y = torch.randint(0, C, (2, 32, 64), generator=g)
y[torch.rand(2, 32, 64, generator=g) < 0.5] = 255
y = y.cuda()
ops.st_loss_fwd(z, t, y, 'ignored')
ops.st_loss_bwd(z, t, y, torch.full((4,), 0.1, device='cuda'), 'ignored')
# round 2: one-pass forward + backward (label pre-pass, fused kernel, checked backward on both of its paths)
for yy in (y, y.to(torch.uint8)):
    sums, counts, used, grad = ops.st_loss_fused(z, t, yy, torch.tensor([1.0, 0.1, 1.0, 0.5], device='cuda'), 'ignored')
    ops.st_loss_bwd_checked(z, t, yy, used.clone(), used, grad, 'ignored')
    ops.st_loss_bwd_checked(z, t, yy, used * 3, used, grad, 'ignored')
pred = torch.randint(0, C, (2, 64, 64), generator=g).cuda()
ops.confusion_matrix(pred, pred.clone(), C)
img = torch.randint(0, 256, (2, 64, 64, 3), dtype=torch.uint8, generator=g).cuda()
lbl = torch.randint(0, C, (2, 64, 64), dtype=torch.uint8, generator=g).cuda()
mask = torch.full((2, 64, 64), 255, dtype=torch.uint8, device='cuda')
ops.copy_paste(img, lbl, mask, img.clone(), lbl.clone(), list(range(14)))
net = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Linear(7, 1031)).cuda()
update_ema_model(net(), net(), 0.99)
# round-1 additions: device PNG writer (fixed + stored segments, odd sizes), nearest resize, validator kernels
import numpy as np  # noqa: E402
rng = np.random.default_rng(0)
for (ph, pw) in ((40, 300), (33, 256), (5, 7)):
    maps = np.stack([np.repeat(np.repeat(rng.integers(0, 19, ((ph + 7) // 8, (pw + 15) // 16)), 8, 0), 16, 1)[:ph, :pw],
                     rng.integers(0, 256, (ph, pw))]).astype(np.uint8)
    ops.PngEncoder(ph, pw, 2).encode_to_host(torch.from_numpy(maps).cuda())
ops.resize_nearest_u8(lbl, (96, 130))
zs = [(torch.randn(2, C, h, w, generator=g) * 3).cuda() for h, w in ((24, 48), (32, 64), (17, 33))]
probs = [ops.softmax_flip_sum(a, a.flip(3).contiguous()) for a in zs]
ops.probs_upsample_argmax(probs, (32, 64))
ops.probs_upsample_argmax(probs[:1], (32, 64))
ops.probs_upsample_argmax(probs + probs[:1], (33, 65))
# general CE (class weights / refer_labels), forward + backward
wts = torch.rand(C, generator=g).cuda()
lab = torch.randint(0, C, (2, 32, 64), generator=g).cuda()
for refer in (None, y):
    sums, cnt = ops.ce_general_fwd(z, lab, wts, refer, 'ignored', 255)
    ops.ce_general_bwd(z, lab, wts, refer, 'ignored', 255, torch.ones(1, device='cuda'))
# PNG writer: both emit variants, many images (offset scan rounds), >32 segments per image
from hiast_b200._lib import lib  # noqa: E402
many = torch.from_numpy(rng.integers(0, 4, (300, 6, 10)).astype(np.uint8)).cuda()
wide = torch.from_numpy(np.repeat(rng.integers(0, 19, (1, 300, 40)), 100, 2).astype(np.uint8)).cuda()
for v in (0, 1):
    lib().hiast_debug_png_variant(v)
    ops.PngEncoder(6, 10, 300).encode_to_host(many)
    ops.PngEncoder(300, 4000, 1).encode_to_host(wide)
lib().hiast_debug_png_variant(1)
# round 2: token ring inside the scan kernel (mailbox looped back to this GPU), window emit with and without the PNG encoder,
# staging ring, writer pool -- through the generator on tiny stride-8 and full-resolution inputs
import ctypes as CT  # noqa: E402
import tempfile  # noqa: E402
from types import SimpleNamespace  # noqa: E402
from hiast_b200.pseudo_label_generator import IASPseudoGenerator  # noqa: E402
box, handle = CT.c_void_p(), (CT.c_ubyte * 64)()
_lib.check(_lib.lib().hiast_ring_create(CT.byref(box), CT.cast(handle, CT.c_void_p)), 'ring_create')
_, _, hist = ops.ias_softmax_hist(logits[:4], B)
st0 = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
ops.ias_threshold_scan(hist[:1].clone(), 1, C, ops.ias_key_lo(C), 0.5, 0.9, 8.0, st0, token=(None, 0, box.value, 3))
ops.ias_threshold_scan(hist[1:].clone(), 1, C, ops.ias_key_lo(C), 0.5, 0.9, 8.0, st0, token=(box.value, 3, None, 0))
torch.cuda.synchronize()
_lib.lib().hiast_ring_destroy(box)
cfg = SimpleNamespace(dataset=SimpleNamespace(num_classes=C),
                      pseudo_policy=SimpleNamespace(type='IAS', batch_size=B, ias=SimpleNamespace(alpha=0.5, beta=0.9, gamma=8.0)),
                      preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=0.99)))


class _LowRes:
    def __call__(self, x):
        return {'logits_lr': x, 'size': (64, 128)}


class _Full:
    def __call__(self, x):
        return {'logits': x}


class _Gen(IASPseudoGenerator):
    def save_data(self):
        pass


class _GenHook(_Gen):
    def save_pseudo_label(self, plbl, img_path):
        pass


host_lr = torch.randn(10, C, 9, 17, generator=g).pin_memory()
host_full = (torch.randn(10, C, 64, 128, generator=g) * 3).pin_memory()
for cls, model, host in ((_Gen, _LowRes(), host_lr), (_GenHook, _LowRes(), host_lr), (_Gen, _Full(), host_full), (_GenHook, _Full(), host_full)):
    loader = [{'images': host[i:i + B], 'image_paths': ['im%d.png' % (i + k) for k in range(B)]} for i in range(0, 10, B)]
    cls(cfg, model=model, loader=loader, save_dir=os.path.join(tempfile.mkdtemp(), 'pl'), window_batches=2).run()
torch.cuda.synchronize()
print('sanitize_small ok')
