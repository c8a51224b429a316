"""Validator on the fused prediction kernels.

Mirrors ``workflows/validator.py`` (reference, /root/reference/code): ``get_multi_scale_and_flip_logits`` :34-55,
``colorize_mask`` / ``save_color_mask`` :57-76 and ``run`` :78-115.  The model and the loader are injected (backbone
and datasets are outside this package), as in ``hiast_b200.pseudo_label_generator``.

What changes: per scale the reference runs softmax, flip, add, interpolate on C x H x W probability tensors, sums the
scales and takes the arg-max -- more than seven full-tensor kernels per scale.  ``predict`` does it in one softmax(+flip)
launch per scale and ONE launch that up-samples every scale, sums and arg-maxes (``hiast_softmax_flip_sum``,
``hiast_probs_upsample_argmax``); the labels are bit-identical to the reference's CUDA path.  The confusion matrix is the
shared-memory bincount of ``hiast_b200.metrics``.
"""

from __future__ import annotations

import os

import numpy as np
import torch
from torch.nn import functional as F

from . import ops
from .metrics import ConfusionMeter

PALETTE_19 = [128, 64, 128, 244, 35, 232, 70, 70, 70, 102, 102, 156, 190, 153, 153, 153, 153, 153, 250, 170, 30,
              220, 220, 0, 107, 142, 35, 152, 251, 152, 70, 130, 180, 220, 20, 60, 255, 0, 0, 0, 0, 142, 0, 0, 70,
              0, 60, 100, 0, 80, 100, 0, 0, 230, 119, 11, 32]
PALETTE_9 = [70, 130, 180, 220, 20, 60, 119, 11, 32, 0, 0, 142, 220, 220, 0, 250, 170, 30, 70, 70, 70, 244, 35, 232,
             128, 64, 128]


class Validator:

    def __init__(self, cfg, model=None, loader=None, device='cuda'):
        self.cfg = cfg
        self.device = torch.device(device)
        if model is None or loader is None:
            raise RuntimeError('hiast_b200.Validator needs model= and loader= (the backbone and the datasets are not '
                               'part of this package; see INTEGRATION.md)')
        self.model = model
        self.v_loader = loader
        path = self.cfg.validate.color_mask_dir_path
        if path is not None:                                                       # :30-32
            assert not os.path.exists(path) or len(os.listdir(path)) == 0
            os.makedirs(path, exist_ok=True)

    # ------------------------------------------------------------------ prediction
    def _scale_probs(self, imgs, size):
        """softmax (+ flipped softmax) of one scale at the scale's own size, one kernel after the forward passes."""
        tmp_imgs = F.interpolate(imgs, size, mode='bilinear', align_corners=True)          # :45 (before the model)
        z0 = self.model(tmp_imgs)['logits'].float().contiguous()
        z1 = None
        if self.cfg.validate.is_flip:                                                      # :48-49
            z1 = self.model(torch.flip(tmp_imgs, dims=[3]))['logits'].float().contiguous()
        return ops.softmax_flip_sum(z0, z1)

    def _check_sizes(self):
        for size in self.cfg.validate.resize_sizes:                                        # :42-43
            assert len(size) == 2 and size[0] <= size[1], \
                'Please input right format of each resize_size: [height, width] and height <= width, such as [512, 1024]'

    def get_multi_scale_and_flip_logits(self, imgs, is_softmax=True):
        """:34-55, same return value ([B,C,H,W] sum over scales).  Kept for callers that want the tensor; ``predict``
        never materialises it."""
        self._check_sizes()
        pred_result_list = []
        for size in self.cfg.validate.resize_sizes:
            if is_softmax:
                pred_result = self._scale_probs(imgs, size)
            else:
                tmp_imgs = F.interpolate(imgs, size, mode='bilinear', align_corners=True)
                pred_result = self.model(tmp_imgs)['logits']
                if self.cfg.validate.is_flip:
                    pred_result = pred_result + torch.flip(self.model(torch.flip(tmp_imgs, dims=[3]))['logits'], dims=[3])
            pred_result_list.append(F.interpolate(pred_result, imgs.size()[2:], mode='bilinear', align_corners=True))
        return sum(pred_result_list)

    def predict(self, imgs):
        """``get_multi_scale_and_flip_logits(imgs).argmax(dim=1)`` (:92-93) as uint8 [B,H,W], fused."""
        self._check_sizes()
        probs = [self._scale_probs(imgs, size) for size in self.cfg.validate.resize_sizes]
        return ops.probs_upsample_argmax(probs, imgs.shape[2:])

    # ------------------------------------------------------------------ colour masks (host, as in the reference)
    def colorize_mask(self, mask):
        """:57-70"""
        from PIL import Image
        if self.cfg.dataset.num_classes == 19:
            palette = PALETTE_19
        elif self.cfg.dataset.num_classes == 9:
            palette = PALETTE_9
        else:
            raise NotImplementedError
        color_mask = Image.fromarray(mask.astype(np.uint8)).convert('P')
        color_mask.putpalette(palette)
        return color_mask

    def save_color_mask(self, lbls_pred, img_paths):
        """:72-76"""
        for lbl_pred, img_path in zip(lbls_pred, img_paths):
            color_mask = self.colorize_mask(lbl_pred)
            color_mask.save(os.path.join(self.cfg.validate.color_mask_dir_path, os.path.basename(img_path)))

    # ------------------------------------------------------------------ run
    def run(self):
        """:78-115.  Returns the dict that the reference prints."""
        print('%% batch_size: {}'.format(getattr(self.cfg.validate, 'batch_size', None)))
        print('%% num_classes: {}'.format(self.cfg.dataset.num_classes))
        print('%% resize_sizes: {}'.format(self.cfg.validate.resize_sizes))
        print('%% is_flip: {}'.format(self.cfg.validate.is_flip))
        print('%% color_mask_dir_path: {}'.format(self.cfg.validate.color_mask_dir_path))
        K = self.cfg.dataset.num_classes
        meter = ConfusionMeter(K, device=self.device)
        if hasattr(self.model, 'eval'):
            self.model.eval()
        with torch.no_grad():
            for data in self.v_loader:
                imgs = data['images'].to(self.device, non_blocking=True)
                lbls = data['labels'].to(self.device, non_blocking=True)
                lbls_pred = self.predict(imgs)
                if lbls.dtype != torch.uint8:                     # labels 0..K-1 / 255: the bincount runs on bytes
                    lbls = lbls.to(torch.uint8)
                meter.update(lbls_pred, lbls.contiguous())
                if self.cfg.validate.color_mask_dir_path is not None:
                    self.save_color_mask(lbls_pred.cpu().numpy(), data['image_paths'])
        synthia = 'SYNTHIA' in self.cfg.dataset.source.type
        res = meter.result(synthia=synthia)
        iou, miou = res['iou'], res['miou']
        if synthia:                                               # :108-113
            print('miou_16: {:.4f}, miou_13: {:.4f}, iou: {}'.format(res['miou_16'], res['miou_13'],
                                                                     {c: round(v, 4) for c, v in enumerate(iou)}))
        else:
            print('miou: {:.4f}, iou: {}'.format(miou, {c: round(v, 4) for c, v in enumerate(iou)}))
        res['confusion_matrix'] = meter.cm
        return res
