"""LOSS registry entries backed by the fused CUDA loss kernels.

Mirrors ``sseg/models/modules/losses.py`` (reference, /root/reference/code): ``ce`` :32-36,
``soft_ce`` :39-41 (+ ``SoftCELoss`` :44-65, ``compute_loss_by_selected_pixel`` :75-89).  Same
signatures ``LOSS[name](logits, labels, weights=None, ignore_index=255, refer_labels=None, region=...)``
returning a 0-d float32 tensor with autograd; the whole chain of log_softmax / mask / product / sum /
count kernels is one forward kernel + one backward kernel (hiast_b200/csrc/loss.cu).

``kl_div`` :16-23 and ``mse`` :9-13 (the other consistency-loss types) run on the
same kernels with a different per-element term.  ``ce`` with class ``weights`` and / or ``refer_labels`` (no HIAST
config reaches it) runs on the plain kernels of csrc/ce_general.cu.  ``BCEWithLogits`` is the adversarial warm-up loss
and out of scope.
"""

from __future__ import annotations

import torch

from . import ops
from ._lib import CST_KLDIV, CST_MSE, CST_SOFTCE, CST_SOFTCE_LOGITS, TERM_CE, TERM_CST, TERM_ENT, TERM_KLD, HiastError
from .registry import LOSS

IGNORE = 255


class FusedSelfTrainingLoss(torch.autograd.Function):
    """out[4] = (CE, KLD, ENT, CST) unweighted; entries of disabled terms are 0.

    Denominators follow the reference: CE / n_conf; KLD / (C*n_conf); ENT / (C*n_ign);
    CST / #non-zero masked products (or / numel when ``cst_mean_all``).  Empty regions give NaN
    (0/0) exactly like the reference.  No host synchronisation anywhere.

    ``hint`` (f32[4] on the device, or None): the upstream gradient the caller EXPECTS for each of the four outputs (its loss
    weights times the expected upstream scalar).  With a hint the forward call already writes the gradient in the same pass
    over (z, t, plbl) (``hiast_st_loss_fused``: 236 B/px instead of 160 + 236); backward compares the scales the real
    upstream gradients imply with the assumed ones ON THE DEVICE and rewrites the gradient only if they differ
    (``hiast_st_loss_bwd_checked``) -- exact in every case, one pass when the expectation holds.
    """

    @staticmethod
    def forward(ctx, z, t, plbl, region, terms, cst_mean_all, hint):
        c = z.shape[1]
        one = None
        if hint is not None and not cst_mean_all and ctx.needs_input_grad[0]:
            one = ops.st_loss_fused(z, t, plbl, hint, region, terms)          # None: configuration not covered
        if one is not None:
            sums, counts, used, grad = one
        else:
            sums, counts = ops.st_loss_fwd(z, t, plbl, region, terms)
            used = grad = None
        cnt = counts.to(torch.float64)
        cst_div = torch.full((), float(z.numel()), dtype=torch.float64, device=z.device) if cst_mean_all else cnt[2]
        denom = torch.stack([cnt[0], c * cnt[0], c * cnt[1], cst_div])
        enabled = _enabled_mask(terms, z.device)
        out = torch.where(enabled, sums / denom, torch.zeros_like(sums)).to(torch.float32)
        ctx.one_pass = one is not None
        if ctx.one_pass:
            ctx.save_for_backward(z, t, plbl, denom, enabled, used, grad)
        else:
            ctx.save_for_backward(z, t, plbl, denom, enabled)
        ctx.region, ctx.terms = region, terms
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.one_pass:
            z, t, plbl, denom, enabled, used, grad = ctx.saved_tensors
        else:
            z, t, plbl, denom, enabled = ctx.saved_tensors
        scales = torch.where(enabled, gout.to(torch.float64) / denom, torch.zeros_like(denom)).to(torch.float32).contiguous()
        if ctx.one_pass:
            grad = ops.st_loss_bwd_checked(z, t, plbl, scales, used, grad, ctx.region, ctx.terms)
        else:
            grad = ops.st_loss_bwd(z, t, plbl, scales, ctx.region, ctx.terms)
        return grad, None, None, None, None, None, None


class FusedTermsLean(torch.autograd.Function):
    """The four loss terms as FOUR outputs (0-d float32 tensors) with everything around the kernels done on the device:
    ``hiast_st_loss_fused_terms`` divides the sums by their counts in its finalize launch, ``hiast_st_loss_bwd_checked_terms``
    derives the gradient scales from the four upstream gradients autograd hands back and records the upstream scalar for
    the next step's expectation.  Forward = 1 tiny multiply (the hint) + 3 chained launches; backward = 1 launch that exits at
    once in the steady state.  (``FusedSelfTrainingLoss`` composes the same numbers from a dozen tiny torch kernels each way;
    it remains the path for every configuration the one-pass kernel does not cover.)  Same bits as that composition."""

    @staticmethod
    def forward(ctx, z, t, plbl, region, terms, hint, weighted):
        # weighted: the outputs are w_k * loss_k with the hint's weights (the products of self_training_segmentor.py:37-52 and
        # their MulBackward nodes move into the two kernels as well)
        one = ops.st_loss_fused_terms(z, t, plbl, hint.tensor(), region, terms, term_weights=hint.weights if weighted else None)
        if one is None:
            raise HiastError('configuration not covered by the one-pass kernel')        # callers check `lean_ok` first
        losses, divisors, used, grad, _, _ = one
        ctx.save_for_backward(z, t, plbl, divisors, used, grad)
        ctx.region, ctx.terms, ctx.hint, ctx.weighted = region, terms, hint, weighted
        ctx.set_materialize_grads(False)
        return tuple(losses.unbind(0))

    @staticmethod
    def backward(ctx, g_ce, g_kld, g_ent, g_cst):
        z, t, plbl, divisors, used, grad = ctx.saved_tensors
        hint = ctx.hint
        gouts = [g if g is None else g.to(torch.float32).contiguous() for g in (g_ce, g_kld, g_ent, g_cst)]
        grad = ops.st_loss_bwd_checked_terms(z, t, plbl, gouts, divisors, used, grad, ctx.region, ctx.terms,
                                             hint_weights=hint.weights, k0=hint.k0,
                                             upstream_out=hint.upstream if gouts[hint.k0] is not None else None,
                                             term_weights=hint.weights if ctx.weighted else None)
        return grad, None, None, None, None, None, None


def lean_ok(z, plbl, terms, cst_mean_all, grad_hint):
    """The conditions under which ``hiast_st_loss_fused_terms`` launches (loss.cu: SoftCE kind, C in {16, 19}, even HW, 8-byte
    aligned tensors) and a gradient is wanted."""
    kind = terms & CST_SOFTCE_LOGITS
    return (grad_hint is not None and not cst_mean_all and kind == CST_SOFTCE and z.requires_grad and torch.is_grad_enabled() and
            z.shape[1] in (16, 19) and z[0, 0].numel() % 2 == 0 and z.shape[0] > 0 and z.data_ptr() % 8 == 0)


def fused_terms_split(logits, plbl, target=None, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT | TERM_CST,
                      grad_hint=None, weighted=False):
    """(CE, KLD, ENT, CST) as four 0-d tensors.  With a ``GradHint`` and a covered configuration: the lean one-pass path; else the
    entries of ``fused_terms``.  ``weighted``: each term comes multiplied by its weight of the ``GradHint`` (required then) --
    ``w * term`` in float32, exactly what the caller would compute."""
    z = _prep_logits(logits)
    t = None
    if terms & TERM_CST:
        t = _prep_logits(target)
        assert t.shape == z.shape                                     # losses.py:50
    y = _prep_labels(plbl)
    if lean_ok(z, y, terms, False, grad_hint) and (t is None or t.data_ptr() % 8 == 0):
        return FusedTermsLean.apply(z, t, y, region, terms, grad_hint, bool(weighted))
    out = fused_terms(z, y, t, region=region, terms=terms, grad_hint=grad_hint)
    if weighted:
        return tuple(w * out[k] for k, w in enumerate(grad_hint.weight_values))
    return out[0], out[1], out[2], out[3]


class GradHint:
    """Expected upstream gradients for ``fused_terms(..., grad_hint=...)``: the caller's four loss weights times the upstream
    scalar of the LAST backward seen on this device (1.0 at first; a loss scaler's factor after one step), kept on the
    device -- nothing here synchronises.  ``hint = float32(upstream) * float32(weight)`` is the arithmetic autograd applies to
    ``weight * loss`` under ``sum(losses).backward()``, so the expectation is met bit for bit whenever upstream / weight
    round-trips in float32 (1.0, powers of two); otherwise the gradient is simply rewritten in backward."""

    _upstream = {}

    def __init__(self, weights, device):
        self.device = torch.device(device)
        self.weight_values = tuple(float(w) for w in weights)
        self.weights = torch.tensor(self.weight_values, dtype=torch.float32, device=self.device)
        key = (self.device.type, self.device.index)
        if key not in GradHint._upstream:
            GradHint._upstream[key] = torch.ones((), dtype=torch.float32, device=self.device)
        self.upstream = GradHint._upstream[key]
        self.k0 = next((k for k, w in enumerate(weights) if float(w) != 0.0), 0)

    def tensor(self):
        return (self.upstream * self.weights).contiguous()

    def observe(self, out):
        """Register a hook on the loss vector ``out``: after its backward, remember upstream = gout[k] / weight[k] of the
        first term with a non-zero weight (two tiny device ops)."""
        if not out.requires_grad:
            return
        wk, up, k0 = self.weights[self.k0], self.upstream, self.k0

        def hook(gout):
            up.copy_(gout[k0] / wk)
            return None
        out.register_hook(hook)


_enabled_cache = {}


def _enabled_mask(terms, device):
    key = (terms & 15, device)
    m = _enabled_cache.get(key)
    if m is None:
        m = torch.tensor([bool(terms & TERM_CE), bool(terms & TERM_KLD), bool(terms & TERM_ENT), bool(terms & TERM_CST)],
                         device=device)
        _enabled_cache[key] = m
    return m


def _prep_logits(logits):
    if not logits.is_cuda:
        raise HiastError('hiast_b200 losses need CUDA tensors (there is no CPU path)')
    if logits.dtype != torch.float32:
        logits = logits.float()
    return logits.contiguous()


def _prep_labels(labels):
    if labels.dtype not in (torch.uint8, torch.int64):
        labels = labels.long()
    return labels.contiguous()


def fused_terms(logits, plbl, target=None, region='ignored', terms=TERM_CE | TERM_KLD | TERM_ENT, cst_mean_all=False,
                grad_hint=None):
    """out[4] = (CE, KLD, ENT, CST).  ``grad_hint`` (a ``GradHint``): take the one-pass forward + backward kernel."""
    z = _prep_logits(logits)
    t = None
    if terms & TERM_CST:
        t = _prep_logits(target)
        assert t.shape == z.shape                                     # losses.py:50
    hint = None
    if grad_hint is not None and z.requires_grad and torch.is_grad_enabled():
        hint = grad_hint.tensor()
    out = FusedSelfTrainingLoss.apply(z, t, _prep_labels(plbl), region, terms, cst_mean_all, hint)
    if hint is not None:
        grad_hint.observe(out)
    return out


class GeneralCE(torch.autograd.Function):
    """CE with class weights and / or a refer_labels region (losses.py:32-36 through :68-89), hiast_ce_general_fwd/bwd."""

    @staticmethod
    def forward(ctx, z, labels, weights, refer_labels, region, ignore_index):
        sums, count = ops.ce_general_fwd(z, labels, weights, refer_labels, region, ignore_index)
        denom = sums[1] if refer_labels is None else count[0].to(torch.float64)
        ctx.save_for_backward(z, labels, weights, refer_labels, denom)
        ctx.region, ctx.ignore_index = region, ignore_index
        return (sums[0] / denom).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        z, labels, weights, refer_labels, denom = ctx.saved_tensors
        scale = (gout.to(torch.float64) / denom).to(torch.float32).reshape(1).contiguous()
        grad = ops.ce_general_bwd(z, labels, weights, refer_labels, ctx.region, ctx.ignore_index, scale)
        return grad, None, None, None, None, None


@LOSS.register('CE')
def ce(logits, labels, weights=None, ignore_index=IGNORE, refer_labels=None, region='confident'):
    """losses.py:32-36.  The hot-path form (no weights, no refer_labels, ignore_index 255) is the CE term of the fused
    kernel; class ``weights`` (f32 [C]) and ``refer_labels`` (+ ``region``) take the general kernels, which keep the
    reference's [B,B,H,W] broadcast of the [B,H,W] loss against the [B,1,H,W] mask (:86-87)."""
    if weights is None and refer_labels is None and ignore_index == IGNORE:
        return fused_terms(logits, labels, terms=TERM_CE)[0]
    if refer_labels is not None and region not in ('ignored', 'confident', 'all'):
        raise ValueError('{} is not a valid region'.format(region))      # losses.py:84
    z = _prep_logits(logits)
    w = None
    if weights is not None:
        w = torch.as_tensor(weights, dtype=torch.float32, device=z.device).contiguous()
        assert w.numel() == z.shape[1]
    r = None if refer_labels is None else _prep_labels(refer_labels)
    return GeneralCE.apply(z, _prep_labels(labels), w, r, region, int(ignore_index))


@LOSS.register('SoftCE')
def soft_ce(logits, labels, weights=None, ignore_index=IGNORE, refer_labels=None, region='confident'):
    """losses.py:39-41.  ``labels`` are soft targets [B,C,H,W] in [0,1] (the reference asserts the range with
    two host syncs, :52; here that check is skipped to stay asynchronous)."""
    if ignore_index != IGNORE:
        raise NotImplementedError('only ignore_index=255 is supported')
    if region not in ('ignored', 'confident', 'all'):
        raise ValueError('{} is not a valid region'.format(region))      # losses.py:84
    if weights is not None:
        assert len(weights) == labels.shape[1]                           # :56
        for c in range(labels.shape[1]):                                 # in-place on the caller's target, like :57-58
            labels[:, c, :, :] *= weights[c]
    if refer_labels is None:
        b, _, h, w = logits.shape
        dummy = torch.zeros((b, h, w), dtype=torch.uint8, device=logits.device)
        return fused_terms(logits, dummy, labels, region='all', terms=TERM_CST, cst_mean_all=True)[3]
    return fused_terms(logits, refer_labels, labels, region=region, terms=TERM_CST)[3]


def _consistency(kind, logits, labels, weights, ignore_index, refer_labels, region):
    if weights is not None:
        import warnings
        warnings.warn('Weights is not available for this loss')              # losses.py:11-12 / :18-19
    if ignore_index != IGNORE:
        raise NotImplementedError('only ignore_index=255 is supported')
    if refer_labels is None:                                                 # reduction='mean' over every element
        b, _, h, w = logits.shape
        dummy = torch.zeros((b, h, w), dtype=torch.uint8, device=logits.device)
        return fused_terms(logits, dummy, labels, region='all', terms=TERM_CST | kind, cst_mean_all=True)[3]
    if region not in ('ignored', 'confident', 'all'):
        raise ValueError('{} is not a valid region'.format(region))
    return fused_terms(logits, refer_labels, labels, region=region, terms=TERM_CST | kind)[3]


@LOSS.register('SoftCE_from_logits')
def soft_ce_from_logits(logits, teacher_logits, weights=None, ignore_index=IGNORE, refer_labels=None, region='confident'):
    """SoftCE whose soft targets are softmax(teacher_logits), with that softmax computed inside the kernel (SURVEY.md
    section 8f rank 4: drops the trainer's F.softmax pass, consistency_self_training_trainer.py:119, 152 B/px)."""
    return _consistency(CST_SOFTCE_LOGITS, logits, teacher_logits, weights, ignore_index, refer_labels, region)


@LOSS.register('MSE')
def mse(logits, labels, weights=None, ignore_index=IGNORE, refer_labels=None, region='ignore'):
    """losses.py:9-13: (logits - labels)^2, mean over all elements or over the non-zero elements of a region."""
    return _consistency(CST_MSE, logits, labels, weights, ignore_index, refer_labels, region)


@LOSS.register('KLDIV')
def kl_div(input_logits, target_logits, weights=None, ignore_index=IGNORE, refer_labels=None, region='confident'):
    """losses.py:16-23: KLDivLoss(log_softmax(input_logits), softmax(target_logits)); both soft-maxes are computed
    inside the kernel.  No gradient flows to ``target_logits`` (the trainer produces them under no_grad)."""
    return _consistency(CST_KLDIV, input_logits, target_logits, weights, ignore_index, refer_labels, region)
