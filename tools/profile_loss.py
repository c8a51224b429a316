#!/usr/bin/env python
"""ncu driver for the fused loss (BASELINE.json configs[2]: 2x19x512x1024, 50 % ignored) and the metric kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiast_b200 import ops  # noqa: E402

z = torch.randn(2, 19, 512, 1024, device='cuda') * 3
t = torch.softmax(torch.randn(2, 19, 512, 1024, device='cuda') * 3, dim=1)
plbl = torch.randint(0, 19, (2, 512, 1024), device='cuda')
plbl[torch.rand(2, 512, 1024, device='cuda') < 0.5] = 255
scales = torch.full((4,), 1e-6, device='cuda')
grad = torch.empty_like(z)
pred = torch.randint(0, 19, (2, 1024, 2048), device='cuda')
tgt = torch.randint(0, 19, (2, 1024, 2048), device='cuda')
img = torch.randint(0, 256, (4, 1024, 2048, 3), dtype=torch.uint8, device='cuda')
lbl = torch.randint(0, 19, (4, 1024, 2048), dtype=torch.uint8, device='cuda')
mask = torch.full((4, 1024, 2048), 255, dtype=torch.uint8, device='cuda')
dlbl = (torch.arange(4 * 1024 * 2048, device='cuda') // 5000 % 19).to(torch.uint8).view(4, 1024, 2048)
gw = torch.tensor([1.0, 0.1, 1.0, 0.5], device='cuda')
for _ in range(3):
    ops.st_loss_fused(z, t, plbl, gw, 'ignored', grad=grad)
    ops.st_loss_fwd(z, t, plbl, 'ignored')
    ops.st_loss_bwd(z, t, plbl, scales, 'ignored', grad=grad)
    ops.confusion_matrix(pred, tgt, 19)
    ops.copy_paste(img, lbl, mask, img.clone(), dlbl, list(range(14)))
torch.cuda.synchronize()
print('ok')
