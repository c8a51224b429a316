"""SelfTrainingSegmentor: the reference's loss composition on top of the fused CUDA loss.

Mirrors ``sseg/models/segmentors/self_training_segmentor.py`` (reference, /root/reference/code):
``SelfTrainingSegmentor`` :9-53, ``build_region_weight`` :128-137, ``_entropy`` :140-150, ``_kld``
:153-163.  ``compute_loss`` returns the same dict (same keys, same order, 0-d tensors with autograd)
that ``workflows/trainer/base_trainer.py:129`` sums.  The reference materialises two [B,C,H,W]
region-weight tensors and runs each term as its own op chain; here all enabled terms of the target
branch come out of ONE forward kernel and their gradient out of ONE backward kernel.

The DeepLabv2 backbone is out of scope (stays stock PyTorch): pass any ``seg_model`` module that
returns ``(logits, backbone)`` like ``sseg/models/modules/seg_models/deeplab_v2.py:58-64``.
"""

from __future__ import annotations

from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F

from ._lib import TERM_CE, TERM_CST, TERM_ENT, TERM_KLD
from .losses import IGNORE, GradHint, fused_terms, fused_terms_split
from .registry import LOSS, MODEL, SEG_MODEL

# What build_region_weight hands to _kld / _entropy instead of a dense [B,C,H,W] tensor.
RegionWeight = namedtuple('RegionWeight', ['plbl', 'region'])


def build_region_weight(t_logits, t_plbl):
    """:128-137.  Returns (confident, ignored) region descriptors; no [B,C,H,W] tensor is built."""
    return RegionWeight(t_plbl, 'confident'), RegionWeight(t_plbl, 'ignored')


def _check_weight(weight, region):
    if not isinstance(weight, RegionWeight):
        raise TypeError('pass the RegionWeight from hiast_b200.segmentor.build_region_weight '
                        '(dense [B,C,H,W] weight tensors are what this path removes)')
    if weight.region != region:
        raise ValueError('this regulariser is defined on the %s region' % region)


def _kld(logits, weight):
    """:153-163  KL(uniform || softmax) regulariser over the confident region."""
    _check_weight(weight, 'confident')
    return fused_terms(logits, weight.plbl, terms=TERM_KLD)[1]


def _entropy(logits, weight):
    """:140-150  entropy regulariser over the ignored region."""
    _check_weight(weight, 'ignored')
    return fused_terms(logits, weight.plbl, terms=TERM_ENT)[2]


@MODEL.register('SelfTrainingSegmentor')
class SelfTrainingSegmentor(nn.Module):

    def __init__(self, cfg, seg_model=None):
        super().__init__()
        self.cfg = cfg
        if seg_model is None:
            # ``build_seg_model(cfg)`` (sseg/models/modules/seg_models/__init__.py:5-8) =
            # SEG_MODEL[cfg.model.seg_model.type](num_classes=..., output_dim=...): the backbone is the host project's
            # (registered there, or brought in by registry.install_into); without one the segmentor only computes losses
            sm = getattr(getattr(cfg, 'model', None), 'seg_model', None)
            if sm is not None and getattr(sm, 'type', None) in SEG_MODEL:
                seg_model = SEG_MODEL[sm.type](num_classes=cfg.dataset.num_classes, output_dim=getattr(sm, 'output_dim', 256))
        self.seg_model = seg_model
        seg_type = cfg.model.predictor.seg_loss.type if hasattr(cfg.model.predictor.seg_loss, 'type') else 'CE'
        self.seg_loss_fun = LOSS[seg_type]
        self.kld_loss_fun = _kld
        self.ent_loss_fun = _entropy
        if cfg.cst_training.is_enabled:
            self.cst_loss_fun = LOSS[cfg.cst_training.cst_loss.type]
        self._fusable = seg_type == 'CE' and (not cfg.cst_training.is_enabled or
                                              cfg.cst_training.cst_loss.type == 'SoftCE')

    def forward(self, t_img):
        """:25-28 (the backbone itself is stock PyTorch and not part of this package)."""
        if self.seg_model is None:
            raise RuntimeError('no seg_model was given: the DeepLabv2 backbone is outside this package')
        t_logits, backbone = self.seg_model(t_img)
        if getattr(self, 'fused_upsample', False) and not torch.is_grad_enabled():
            # pseudo-labelling: hand the stride-8 logits to the generator, which fuses the up-sampling into phase A
            return {'logits_lr': t_logits, 'size': tuple(t_img.shape[2:]), 'backbone': backbone}
        t_logits = F.interpolate(t_logits, size=t_img.shape[2:], mode='bilinear', align_corners=True)
        return {'logits': t_logits, 'backbone': backbone}

    def compute_loss(self, t_logits, t_plbl, t_cst_lbl=None, s_logits=None, s_lbl=None):
        """:30-53."""
        cfg = self.cfg
        losses = {}
        if s_lbl is not None:
            losses['source_seg_loss'] = self.seg_loss_fun(s_logits, s_lbl)
        w_seg = cfg.model.predictor.seg_loss.target_pseudo_weight
        w_kld = cfg.model.predictor.kld_loss.weight
        w_ent = cfg.model.predictor.ent_loss.weight
        use_cst = t_cst_lbl is not None and cfg.cst_training.is_enabled and cfg.cst_training.cst_loss.weight > 0
        if not self._fusable:
            # other registered loss types: compose term by term like the reference
            losses['target_seg_loss'] = w_seg * self.seg_loss_fun(t_logits, t_plbl)
            w_conf, w_ign = build_region_weight(t_logits, t_plbl)
            if w_kld > 0:
                losses['kld_confident_loss'] = w_kld * self.kld_loss_fun(t_logits, w_conf)
            if w_ent > 0:
                losses['ent_ignored_loss'] = w_ent * self.ent_loss_fun(t_logits, w_ign)
            if use_cst:
                losses['cst_loss'] = cfg.cst_training.cst_loss.weight * self.cst_loss_fun(
                    t_logits, t_cst_lbl, refer_labels=t_plbl, region=cfg.cst_training.cst_loss.region)
            return losses
        terms = TERM_CE | (TERM_KLD if w_kld > 0 else 0) | (TERM_ENT if w_ent > 0 else 0) | (TERM_CST if use_cst else 0)
        region = cfg.cst_training.cst_loss.region if use_cst else 'ignored'
        if region not in ('ignored', 'confident', 'all'):
            raise ValueError('{} is not a valid region'.format(region))
        # one pass over (z, t, plbl) for forward AND backward: the gradient is written for the upstream gradients this
        # composition produces under sum(losses).backward() (base_trainer.py:129-133) and verified / redone in backward
        weights = (w_seg, w_kld if w_kld > 0 else 0.0, w_ent if w_ent > 0 else 0.0,
                   cfg.cst_training.cst_loss.weight if use_cst else 0.0)
        hint = None
        if getattr(self, 'one_pass', True) and t_logits.is_cuda:
            key = (weights, t_logits.device)
            hint = self.__dict__.setdefault('_hints', {}).get(key)
            if hint is None:
                hint = self._hints[key] = GradHint(weights, t_logits.device)
        if hint is not None:
            # the weighted terms themselves come out of the kernels (w * loss in float32, as the products below would give)
            out = fused_terms_split(t_logits, t_plbl, t_cst_lbl if use_cst else None, region=region, terms=terms, grad_hint=hint,
                                    weighted=True)
            scale = (1.0, 1.0, 1.0, 1.0)
        else:
            out = fused_terms_split(t_logits, t_plbl, t_cst_lbl if use_cst else None, region=region, terms=terms)
            scale = weights
        losses['target_seg_loss'] = out[0] if hint is not None else scale[0] * out[0]
        if w_kld > 0:
            losses['kld_confident_loss'] = out[1] if hint is not None else scale[1] * out[1]
        if w_ent > 0:
            losses['ent_ignored_loss'] = out[2] if hint is not None else scale[2] * out[2]
        if use_cst:
            losses['cst_loss'] = out[3] if hint is not None else scale[3] * out[3]
        return losses
