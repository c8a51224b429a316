"""Oracle: region-adaptive regularisation + consistency losses (torch, any device).

Restates (reference, /root/reference/code):
``sseg/models/segmentors/self_training_segmentor.py`` -- ``compute_loss`` :30-53,
``build_region_weight`` :128-137, ``_entropy`` :140-150, ``_kld`` :153-163 -- and
``sseg/models/modules/losses.py`` -- ``ce`` :32-36, ``soft_ce`` :39-41,
``SoftCELoss`` :44-65, ``compute_loss_by_selected_pixel`` :75-89.

The op sequence (full [B,C,H,W] weight tensors, log_softmax, masked multiply,
sum, divide by an element count) is kept so that fp32 results track the
reference's own autograd path as closely as torch allows; gradients come from
torch autograd on these expressions.  Test infrastructure only.
"""

from __future__ import annotations

import torch
from torch.nn import functional as F

IGNORE = 255


def region_weights(logits, plbl):
    """(w_confident, w_ignored), each [B,C,H,W] like logits  (:128-137)."""
    valid = (plbl != IGNORE).to(logits.dtype).unsqueeze(1)      # [B,1,H,W]
    ones = torch.ones_like(logits)
    return ones * valid, ones * (1 - valid)


def entropy_reg(logits, weight):
    """:140-150 -- sum(-p * w * logp) / #(w > 0)."""
    n = int((weight > 0).sum().item())
    logp = torch.log_softmax(logits, dim=1)
    ent = -torch.softmax(logits, dim=1) * weight * logp
    return torch.sum(ent) / n


def kld_reg(logits, weight):
    """:153-163 -- sum(-(1/C) * w * logp) / #(w > 0)."""
    n = int((weight > 0).sum().item())
    logp = torch.log_softmax(logits, dim=1)
    c = logits.size(1)
    return torch.sum(-1 / c * weight * logp) / n


def ce(logits, labels, ignore_index=IGNORE):
    """losses.py:32-36 with refer_labels=None -> nn.CrossEntropyLoss(ignore_index) mean."""
    return F.cross_entropy(logits, labels, ignore_index=ignore_index)


def ce_general(logits, labels, weights=None, ignore_index=IGNORE, refer_labels=None, region='confident'):
    """losses.py:32-36 with class weights and / or refer_labels (:68-72, :75-89).  With refer_labels the [B,H,W] loss times
    the [B,1,H,W] mask broadcasts to [B,B,H,W] exactly as in the reference."""
    if refer_labels is None:
        return F.cross_entropy(logits, labels, weight=weights, ignore_index=ignore_index)
    loss_tensor = F.cross_entropy(logits, labels, weight=weights, reduction='none')
    loss_tensor = loss_tensor * _region_mask(refer_labels, ignore_index, region).unsqueeze(dim=1)
    return loss_tensor.sum() / (loss_tensor != 0).sum()


def _region_mask(refer_labels, ignore_index, region):
    if region == 'ignored':
        return refer_labels == ignore_index
    if region == 'confident':
        return refer_labels != ignore_index
    if region == 'all':
        return torch.ones_like(refer_labels, dtype=torch.bool)
    raise ValueError('{} is not a valid region'.format(region))


def soft_ce(logits, target, refer_labels=None, region='confident', ignore_index=IGNORE):
    """losses.py:39-41 -> :44-61 -> :68-72 -> :75-89."""
    assert logits.shape == target.shape
    assert target.min().item() >= 0 and target.max().item() <= 1
    nll = -F.log_softmax(logits, dim=1)
    if refer_labels is None:
        return (nll * target).sum() / target.numel()
    per_elem = (nll * target) * _region_mask(refer_labels, ignore_index, region).unsqueeze(1)
    return per_elem.sum() / (per_elem != 0).sum()


def _by_region(per_elem, numel, refer_labels, region, ignore_index):
    """losses.py:68-72 + :75-89 for a reduction='none' tensor."""
    if refer_labels is None:
        return per_elem.sum() / numel
    per_elem = per_elem * _region_mask(refer_labels, ignore_index, region).unsqueeze(1)
    return per_elem.sum() / (per_elem != 0).sum()


def mse(logits, labels, refer_labels=None, region='ignore', ignore_index=IGNORE):
    """losses.py:9-13 (nn.MSELoss / nn.MSELoss(reduction='none'))."""
    return _by_region(F.mse_loss(logits, labels, reduction='none'), logits.numel(), refer_labels, region, ignore_index)


def kl_div(input_logits, target_logits, refer_labels=None, region='confident', ignore_index=IGNORE):
    """losses.py:16-23 (nn.KLDivLoss default 'mean' = mean over all elements / reduction='none')."""
    inp = F.log_softmax(input_logits, dim=1)
    tgt = F.softmax(target_logits, dim=1)
    return _by_region(F.kl_div(inp, tgt, reduction='none'), inp.numel(), refer_labels, region, ignore_index)


def compute_loss(t_logits, t_plbl, t_cst_lbl=None, s_logits=None, s_lbl=None, *,
                 w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, cst_region='ignored',
                 cst_enabled=True):
    """SelfTrainingSegmentor.compute_loss :30-53 (seg loss type CE, cst loss type SoftCE)."""
    out = {}
    if s_lbl is not None:
        out['source_seg_loss'] = ce(s_logits, s_lbl)
    out['target_seg_loss'] = w_seg * ce(t_logits, t_plbl)
    w_conf, w_ign = region_weights(t_logits, t_plbl)
    if w_kld > 0:
        out['kld_confident_loss'] = w_kld * kld_reg(t_logits, w_conf)
    if w_ent > 0:
        out['ent_ignored_loss'] = w_ent * entropy_reg(t_logits, w_ign)
    if t_cst_lbl is not None and cst_enabled and w_cst > 0:
        out['cst_loss'] = w_cst * soft_ce(t_logits, t_cst_lbl, refer_labels=t_plbl, region=cst_region)
    return out
