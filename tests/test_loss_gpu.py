"""GPU parity of the fused loss kernels and their reference-facing wrappers (1e-5 relative)."""

import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import losses as oloss

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL = 1e-5      # BASELINE.json north_star: losses and gradients within 1e-5 relative


def make_cfg(spec):
    return SimpleNamespace(
        model=SimpleNamespace(predictor=SimpleNamespace(
            seg_loss=SimpleNamespace(type='CE', target_pseudo_weight=spec['w_seg']),
            kld_loss=SimpleNamespace(weight=spec['w_kld']),
            ent_loss=SimpleNamespace(weight=spec['w_ent']))),
        cst_training=SimpleNamespace(is_enabled=True, cst_loss=SimpleNamespace(
            type='SoftCE', weight=spec['w_cst'], region=spec['region'])))


def grad_close(got, want):
    got, want = got.double(), want.double()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= RTOL * scale + 1e-12


@pytest.mark.parametrize('name', list(gi.LOSS_SPECS))
def test_compute_loss_vs_reference_fixture(name):
    """SelfTrainingSegmentor.compute_loss: same keys, values and gradient as the reference's own run."""
    from hiast_b200.segmentor import SelfTrainingSegmentor
    spec = gi.LOSS_SPECS[name]
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    z = torch.from_numpy(gold['z']).cuda().requires_grad_(True)
    t = torch.from_numpy(gold['t']).cuda()
    plbl = torch.from_numpy(gold['plbl']).cuda()
    s_z = s_lbl = None
    if spec['source']:
        s_z = torch.from_numpy(gold['s_z']).cuda().requires_grad_(True)
        s_lbl = torch.from_numpy(gold['s_lbl']).cuda()
    seg = SelfTrainingSegmentor(make_cfg(spec))
    out = seg.compute_loss(z, plbl, t, s_z, s_lbl)
    assert list(out.keys()) == list(gold['keys'])
    for v in out.values():
        assert v.dim() == 0 and v.dtype == torch.float32
    np.testing.assert_allclose([v.item() for v in out.values()], gold['values'], rtol=RTOL)
    sum(v.mean() for v in out.values()).backward()                  # base_trainer.py:129
    grad_close(z.grad.cpu(), torch.from_numpy(gold['grad']))
    if spec['source']:
        grad_close(s_z.grad.cpu(), torch.from_numpy(gold['s_grad']))


@pytest.mark.parametrize('region', ['ignored', 'confident', 'all'])
@pytest.mark.parametrize('shape', [(2, 19, 64, 128), (1, 16, 33, 52), (2, 7, 15, 21)])
def test_fused_terms_vs_oracle_on_cuda(shape, region):
    """Vector (C=19/16) and generic (C=7, odd HW) kernels against the oracle's torch expressions on CUDA."""
    from hiast_b200.losses import fused_terms
    b, c, h, w = shape
    g = torch.Generator().manual_seed(b * 1000 + c + h)
    z = (torch.randn(b, c, h, w, generator=g) * 3).cuda().requires_grad_(True)
    t = torch.softmax(torch.randn(b, c, h, w, generator=g) * 3, dim=1).cuda()
    plbl = torch.randint(0, c, (b, h, w), generator=g)
    plbl[torch.rand(b, h, w, generator=g) < 0.4] = 255
    plbl = plbl.cuda()
    wts = torch.tensor([1.0, 0.1, 1.0, 0.5], device='cuda')
    for lbl in (plbl, plbl.to(torch.uint8)):                          # int64 and uint8 label inputs
        z.grad = None
        out = fused_terms(z, lbl, t, region=region, terms=15)
        (out * wts).sum().backward()
        z2 = z.detach().clone().requires_grad_(True)
        ref = oloss.compute_loss(z2, plbl, t, cst_region=region)
        sum(ref.values()).backward()
        np.testing.assert_allclose((out * wts).tolist(), [v.item() for v in ref.values()], rtol=RTOL)
        grad_close(z.grad, z2.grad)


def test_config3_full_size():
    """BASELINE.json configs[2]: 2x19x512x1024, 50 % ignored, CE 1.0 / KLD 0.1 / ENT 1.0 / SoftCE 0.5 'ignored'."""
    from hiast_b200.segmentor import SelfTrainingSegmentor
    spec = dict(w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, region='ignored')
    z = (torch.randn(2, 19, 512, 1024, generator=torch.Generator().manual_seed(0)) * 3).cuda().requires_grad_(True)
    t = torch.softmax(torch.randn(2, 19, 512, 1024, generator=torch.Generator().manual_seed(1)) * 3, dim=1).cuda()
    g2 = torch.Generator().manual_seed(2)
    plbl = torch.randint(0, 19, (2, 512, 1024), generator=g2)
    plbl[torch.rand(2, 512, 1024, generator=g2) < 0.5] = 255
    plbl = plbl.cuda()
    out = SelfTrainingSegmentor(make_cfg(spec)).compute_loss(z, plbl, t)
    sum(v.mean() for v in out.values()).backward()
    z2 = z.detach().clone().requires_grad_(True)
    ref = oloss.compute_loss(z2, plbl, t)
    sum(v.mean() for v in ref.values()).backward()
    np.testing.assert_allclose([v.item() for v in out.values()], [v.item() for v in ref.values()], rtol=RTOL)
    grad_close(z.grad, z2.grad)
    # determinism of the two-stage reduction
    out2 = SelfTrainingSegmentor(make_cfg(spec)).compute_loss(z.detach(), plbl, t)
    assert [v.item() for v in out2.values()] == [v.item() for v in out.values()]


def test_registry_losses():
    """LOSS['CE'] / LOSS['SoftCE'] keep the reference signatures (losses.py:32-41)."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import LOSS
    g = torch.Generator().manual_seed(5)
    z = (torch.randn(2, 19, 16, 32, generator=g) * 2).cuda().requires_grad_(True)
    t = torch.softmax(torch.randn(2, 19, 16, 32, generator=g), dim=1).cuda()
    y = torch.randint(0, 19, (2, 16, 32), generator=g)
    y[torch.rand(2, 16, 32, generator=g) < 0.3] = 255
    y = y.cuda()
    np.testing.assert_allclose(LOSS['CE'](z, y).item(), oloss.ce(z, y).item(), rtol=RTOL)
    for region in ('ignored', 'confident', 'all'):
        got = LOSS['SoftCE'](z, t, refer_labels=y, region=region)
        np.testing.assert_allclose(got.item(), oloss.soft_ce(z, t, refer_labels=y, region=region).item(), rtol=RTOL)
    np.testing.assert_allclose(LOSS['SoftCE'](z, t).item(), oloss.soft_ce(z, t).item(), rtol=RTOL)   # refer_labels=None
    with pytest.raises(ValueError):
        LOSS['SoftCE'](z, t, refer_labels=y, region='bogus')


def test_softce_divisor_counts_nonzero_products():
    """losses.py:89 divides by the number of non-zero masked products: zero targets / exact log-softmax zeros
    shrink the divisor."""
    from hiast_b200 import ops
    g = torch.Generator().manual_seed(9)
    z = (torch.randn(1, 19, 8, 16, generator=g) * 3)
    z[0, 4, :, :8] += 60.0                      # log_softmax == 0 exactly for class 4 there
    t = torch.softmax(torch.randn(1, 19, 8, 16, generator=g), dim=1)
    t[0, 7] = 0.0                               # zero targets
    y = torch.full((1, 8, 16), 255, dtype=torch.int64)
    z, t, y = z.cuda(), t.cuda(), y.cuda()
    sums, counts = ops.st_loss_fwd(z, t, y, 'ignored')
    ref = (-torch.log_softmax(z, dim=1) * t)
    assert counts[2].item() == int((ref != 0).sum().item()) < 19 * 8 * 16
    np.testing.assert_allclose(sums[3].item() / counts[2].item(), oloss.soft_ce(z, t, refer_labels=y, region='ignored').item(),
                               rtol=RTOL)


def test_empty_regions_give_nan_like_the_reference():
    from hiast_b200.losses import fused_terms
    z = torch.randn(1, 19, 4, 8).cuda().requires_grad_(True)
    y = torch.zeros((1, 4, 8), dtype=torch.int64).cuda()          # nothing ignored
    out = fused_terms(z, y, None, terms=7)
    assert torch.isfinite(out[0]) and torch.isfinite(out[1]) and torch.isnan(out[2])     # entropy: 0/0
    out.sum().backward()
    assert torch.isnan(z.grad).all()                                # inf * 0 in the reference's autograd


def test_kldiv_and_mse_consistency_variants_vs_reference_fixture():
    """LOSS['KLDIV'] / LOSS['MSE'] (losses.py:9-23): values and gradients of the reference's own run, all regions,
    vector (C=19) kernels; then the generic kernels (C=7, odd size) against the oracle on CUDA."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import LOSS
    gold = np.load(os.path.join(GOLD, 'loss_cst_variants.npz'))
    z0, t, tz, plbl = (torch.from_numpy(gold[k]).cuda() for k in ('z', 't', 'tz', 'plbl'))
    for kind, name, tgt in (('kldiv', 'KLDIV', tz), ('mse', 'MSE', t)):
        for region in ('none', 'ignored', 'confident', 'all'):
            z = z0.clone().requires_grad_(True)
            val = LOSS[name](z, tgt) if region == 'none' else LOSS[name](z, tgt, refer_labels=plbl, region=region)
            val.backward()
            np.testing.assert_allclose(val.item(), gold['%s_%s' % (kind, region)], rtol=RTOL)
            grad_close(z.grad.cpu(), torch.from_numpy(gold['%s_%s_grad' % (kind, region)]))
    g = torch.Generator().manual_seed(21)
    z0 = (torch.randn(2, 7, 9, 13, generator=g) * 2).cuda()
    tz = (torch.randn(2, 7, 9, 13, generator=g) * 2).cuda()
    y = torch.randint(0, 7, (2, 9, 13), generator=g)
    y[torch.rand(2, 9, 13, generator=g) < 0.5] = 255
    y = y.cuda()
    for name, ofn in (('KLDIV', oloss.kl_div), ('MSE', oloss.mse)):
        z = z0.clone().requires_grad_(True)
        LOSS[name](z, tz, refer_labels=y, region='ignored').backward()
        z2 = z0.clone().requires_grad_(True)
        ofn(z2, tz, refer_labels=y, region='ignored').backward()
        grad_close(z.grad, z2.grad)


def test_segmentor_with_kldiv_consistency():
    """cst_loss.type = KLDIV goes through the per-term composition of compute_loss (self_training_segmentor.py:49-51)."""
    from hiast_b200.segmentor import SelfTrainingSegmentor
    spec = dict(w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, region='confident')
    cfg = make_cfg(spec)
    cfg.cst_training.cst_loss.type = 'KLDIV'
    g = torch.Generator().manual_seed(33)
    z = (torch.randn(2, 19, 16, 24, generator=g) * 3).cuda().requires_grad_(True)
    tz = (torch.randn(2, 19, 16, 24, generator=g) * 3).cuda()
    y = torch.randint(0, 19, (2, 16, 24), generator=g)
    y[torch.rand(2, 16, 24, generator=g) < 0.5] = 255
    y = y.cuda()
    out = SelfTrainingSegmentor(cfg).compute_loss(z, y, tz)
    assert list(out) == ['target_seg_loss', 'kld_confident_loss', 'ent_ignored_loss', 'cst_loss']
    sum(v.mean() for v in out.values()).backward()
    z2 = z.detach().clone().requires_grad_(True)
    w_conf, w_ign = oloss.region_weights(z2, y)
    ref = [oloss.ce(z2, y), 0.1 * oloss.kld_reg(z2, w_conf), oloss.entropy_reg(z2, w_ign),
           0.5 * oloss.kl_div(z2, tz, refer_labels=y, region='confident')]
    sum(ref).backward()
    np.testing.assert_allclose([v.item() for v in out.values()], [v.item() for v in ref], rtol=RTOL)
    grad_close(z.grad, z2.grad)


def test_softce_from_teacher_logits_equals_softce_of_softmax():
    """8f rank 4: the teacher softmax fused into the kernel gives the same loss / gradient as feeding probabilities."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import LOSS
    g = torch.Generator().manual_seed(44)
    for shape in [(2, 19, 16, 24), (1, 7, 9, 11)]:
        z0 = (torch.randn(*shape, generator=g) * 3).cuda()
        tz = (torch.randn(*shape, generator=g) * 3).cuda()
        y = torch.randint(0, shape[1], (shape[0],) + shape[2:], generator=g)
        y[torch.rand(y.shape, generator=g) < 0.5] = 255
        y = y.cuda()
        z1 = z0.clone().requires_grad_(True)
        a = LOSS['SoftCE_from_logits'](z1, tz, refer_labels=y, region='ignored')
        a.backward()
        z2 = z0.clone().requires_grad_(True)
        b = oloss.soft_ce(z2, torch.softmax(tz, dim=1), refer_labels=y, region='ignored')
        b.backward()
        np.testing.assert_allclose(a.item(), b.item(), rtol=RTOL)
        grad_close(z1.grad, z2.grad)


@pytest.mark.parametrize('region', ['ignored', 'confident', 'all'])
def test_packed_pair_kernels_equal_scalar_kernels(region):
    """The f32x2 kernels (SoftCE kind) against the scalar vector kernels: integer counts identical, sums / gradients to
    float rounding.  Teacher probabilities with exact zeros exercise the recount of the non-zero-product divisor."""
    from hiast_b200 import _lib, ops
    g = torch.Generator().manual_seed(11)
    z = (torch.randn(2, 19, 96, 160, generator=g) * 3).cuda()
    z[0, 3, :8] = -1e4                                       # expf underflow lanes
    t = torch.softmax(torch.randn(2, 19, 96, 160, generator=g) * 40, dim=1).cuda()      # many exact zeros
    plbl = torch.randint(0, 19, (2, 96, 160), generator=g)
    plbl[torch.rand(2, 96, 160, generator=g) < 0.5] = 255
    plbl = plbl.cuda()
    scales = torch.tensor([0.3, 0.02, 0.7, 0.11], device='cuda')
    res = {}
    for scalar in (1, 0):
        _lib.lib().hiast_debug_loss_scalar(scalar)
        try:
            sums, counts = ops.st_loss_fwd(z, t, plbl, region)
            grad = ops.st_loss_bwd(z, t, plbl, scales, region)
        finally:
            _lib.lib().hiast_debug_loss_scalar(0)
        res[scalar] = (sums.cpu().numpy(), counts.cpu().numpy(), grad)
    assert np.array_equal(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-6)
    assert int(res[0][1][2]) < (plbl.numel() * 19)           # zeros were really present
    grad_close(res[0][2], res[1][2])


# ------------------------------------------------------------------ CE with class weights / refer_labels (losses.py:32-36)
CE_GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'loss_ce_general.npz')


@pytest.mark.parametrize('key', list(gi.CE_GENERAL_SPECS))
def test_ce_general_matches_reference_fixture_and_oracle(key):
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import LOSS
    from oracle import losses as oloss
    gold = np.load(CE_GOLD)
    spec = gi.CE_GENERAL_SPECS[key]
    z, labels, weights, refer = gi.ce_general_inputs(spec)
    for case, kw in gi.ce_general_cases(labels, weights, refer).items():
        kw_dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
        zz = z.cuda().requires_grad_(True)
        val = LOSS['CE'](zz, **kw_dev)
        val.backward()
        want, want_grad = gold['%s_%s' % (key, case)], gold['%s_%s_grad' % (key, case)]
        np.testing.assert_allclose(val.item(), want, rtol=1e-5, err_msg=case)
        np.testing.assert_allclose(zz.grad.cpu().numpy(), want_grad, rtol=1e-4, atol=1e-8, err_msg=case)
        zo = z.cuda().requires_grad_(True)                       # the reference expression on CUDA
        ref = oloss.ce_general(zo, **kw_dev)
        ref.backward()
        np.testing.assert_allclose(val.item(), ref.item(), rtol=1e-5, err_msg=case)
        np.testing.assert_allclose(zz.grad.cpu().numpy(), zo.grad.cpu().numpy(), rtol=1e-4, atol=1e-8, err_msg=case)
    # uint8 labels take the same path
    lbl8 = kw_dev['labels'].to(torch.uint8)
    a = LOSS['CE'](z.cuda(), lbl8, weights=weights.cuda())
    b = LOSS['CE'](z.cuda(), kw_dev['labels'], weights=weights.cuda())
    assert a.item() == b.item()
    with pytest.raises(ValueError):
        LOSS['CE'](z.cuda(), labels.cuda(), refer_labels=refer.cuda(), region='nowhere')


# ------------------------------------------------------------------ one-pass forward + backward (hiast_st_loss_fused)
def _loss_inputs(shape, seed, ignore_frac=0.5, uint8=False):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(seed)
    z = (torch.randn(b, c, h, w, generator=g) * 3).cuda()
    t = torch.softmax(torch.randn(b, c, h, w, generator=g) * 3, dim=1).cuda()
    y = torch.randint(0, c, (b, h, w), generator=g)
    y[torch.rand(b, h, w, generator=g) < ignore_frac] = 255
    return z, t, (y.to(torch.uint8) if uint8 else y).cuda()


@pytest.mark.parametrize('region', ['ignored', 'confident', 'all'])
@pytest.mark.parametrize('shape,uint8', [((2, 19, 64, 128), False), ((1, 16, 33, 52), True), ((3, 19, 17, 26), False)])
def test_one_pass_kernel_equals_two_pass_kernels(region, shape, uint8):
    """hiast_st_loss_fused == hiast_st_loss_fwd + hiast_st_loss_bwd with the scales it reports, and
    hiast_st_loss_bwd_checked leaves the gradient alone for equal scales / rewrites it for different ones."""
    from hiast_b200 import ops
    z, t, y = _loss_inputs(shape, sum(shape))
    y = y.to(torch.uint8) if uint8 else y
    gw = torch.tensor([1.0, 0.1, 1.0, 0.5], device='cuda')
    res = ops.st_loss_fused(z, t, y, gw, region)
    assert res is not None
    sums, counts, used, grad = res
    sums2, counts2 = ops.st_loss_fwd(z, t, y, region)
    assert torch.equal(counts, counts2)
    np.testing.assert_allclose(sums.cpu().numpy(), sums2.cpu().numpy(), rtol=1e-12)
    c = shape[1]
    cnt = counts.double()
    want_scales = (gw.double() / torch.stack([cnt[0], c * cnt[0], c * cnt[1], cnt[2]])).float()
    assert torch.equal(used, want_scales)                          # counts[2] == C * n_region here (no zero product)
    want = ops.st_loss_bwd(z, t, y, used, region)
    assert torch.equal(grad, want)
    # the check: same scales -> untouched (sentinel survives); other scales -> rewritten
    sentinel = torch.full_like(grad, 7.0)
    ops.st_loss_bwd_checked(z, t, y, used.clone(), used, sentinel, region)
    assert bool((sentinel == 7.0).all())
    ops.st_loss_bwd_checked(z, t, y, used * 2, used, sentinel, region)
    assert torch.equal(sentinel, ops.st_loss_bwd(z, t, y, used * 2, region))


def test_one_pass_through_compute_loss_adapts_to_the_upstream_scale():
    """SelfTrainingSegmentor.compute_loss takes the one-pass kernel; results equal the two-pass composition for upstream
    1.0 (expectation met), for a loss-scaled backward (first step: gradient rewritten; later steps: expectation adapted)
    and when a SoftCE product is exactly zero (divisor != C * n_ign: rewritten)."""
    from hiast_b200.segmentor import SelfTrainingSegmentor
    spec = dict(w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, region='ignored')
    z0, t, y = _loss_inputs((2, 19, 48, 64), 5)
    t[:, 3] = 0.0                                                  # exact zeros among the products of every ignored pixel
    one, two = SelfTrainingSegmentor(make_cfg(spec)), SelfTrainingSegmentor(make_cfg(spec))
    two.one_pass = False
    for scale in (1.0, 1.0, 1024.0, 1024.0, 1024.0, 3.7, 1.0):
        grads, vals = [], []
        for seg in (one, two):
            z = z0.clone().requires_grad_(True)
            out = seg.compute_loss(z, y, t)
            (sum(out.values()) * scale).backward()
            grads.append(z.grad)
            vals.append([v.item() for v in out.values()])
        np.testing.assert_allclose(vals[0], vals[1], rtol=1e-7)
        assert torch.equal(grads[0], grads[1]), scale
    # without exact zeros and with the expected upstream the backward is a no-op on the forward's gradient
    z0, t, y = _loss_inputs((2, 19, 48, 64), 6)
    z = z0.clone().requires_grad_(True)
    out = one.compute_loss(z, y, t)
    sum(out.values()).backward()
    z2 = z0.clone().requires_grad_(True)
    out2 = two.compute_loss(z2, y, t)
    sum(out2.values()).backward()
    assert torch.equal(z.grad, z2.grad)


def test_one_pass_empty_regions_reproduce_the_reference_nans():
    from hiast_b200.segmentor import SelfTrainingSegmentor
    spec = dict(w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, region='ignored')
    z0, t, y = _loss_inputs((1, 19, 16, 32), 9, ignore_frac=0.0)   # no ignored pixel: ENT and CST are 0/0
    seg = SelfTrainingSegmentor(make_cfg(spec))
    z = z0.clone().requires_grad_(True)
    out = seg.compute_loss(z, y, t)
    assert torch.isnan(out['ent_ignored_loss']) and torch.isnan(out['cst_loss'])
    assert torch.isfinite(out['target_seg_loss'])
    sum(out.values()).backward()
    assert torch.isnan(z.grad).all()


# ------------------------------------------------------------------ lean one-pass path (hiast_st_loss_fused_terms / _bwd_checked_terms)
@pytest.mark.parametrize('region', ['ignored', 'confident', 'all'])
def test_fused_terms_kernel_outputs_equal_the_host_composition(region):
    """losses / divisors written by the finalize launch == sums / counts composed with torch ops, bit for bit; the scales the
    backward kernel derives from four upstream gradients == the host arithmetic (gradient identical to the checked call)."""
    from hiast_b200 import ops
    z, t, y = _loss_inputs((2, 19, 40, 64), 11)
    gw = torch.tensor([1.0, 0.1, 1.0, 0.5], device='cuda')
    losses, divisors, used, grad, sums, counts = ops.st_loss_fused_terms(z, t, y, gw, region)
    sums2, counts2, used2, grad2 = ops.st_loss_fused(z, t, y, gw, region)
    assert torch.equal(sums, sums2) and torch.equal(counts, counts2) and torch.equal(used, used2) and torch.equal(grad, grad2)
    cnt = counts.double()
    want_div = torch.stack([cnt[0], 19 * cnt[0], 19 * cnt[1], cnt[2]])
    assert torch.equal(divisors, want_div)
    assert torch.equal(losses, (sums / want_div).float())
    # backward: upstream gradients that differ from the assumption -> rewritten with the scales the host would compute
    gouts = [torch.tensor(v, device='cuda') for v in (3.0, 0.25, 1.5, 0.75)]
    want_scales = (torch.stack(gouts).double() / want_div).float()
    want = ops.st_loss_bwd(z, t, y, want_scales, region)
    up = torch.zeros((), device='cuda')
    got = ops.st_loss_bwd_checked_terms(z, t, y, gouts, divisors, used, grad.clone(), region, hint_weights=gw, k0=1, upstream_out=up)
    assert torch.equal(got, want)
    assert up.item() == np.float32(0.25) / np.float32(0.1)
    # a term without upstream gradient counts as scale 0; equal scales leave the gradient alone
    gouts[2] = None
    want_scales[2] = 0.0
    got = ops.st_loss_bwd_checked_terms(z, t, y, gouts, divisors, used, grad.clone(), region)
    assert torch.equal(got, ops.st_loss_bwd(z, t, y, want_scales, region))
    sentinel = torch.full_like(grad, 7.0)
    ops.st_loss_bwd_checked_terms(z, t, y, [g for g in gw], divisors, used, sentinel, region)
    assert bool((sentinel == 7.0).all())


def test_lean_path_is_taken_and_equals_the_composed_path():
    """compute_loss goes through FusedTermsLean (four outputs, no select / stack kernels) and gives the same values and the same
    gradient bits as the composed one-pass path; an unused term (no gradient flows into it) is handled."""
    from hiast_b200 import losses as L
    z0, t, y = _loss_inputs((2, 19, 48, 64), 21)
    hint = L.GradHint((1.0, 0.1, 1.0, 0.5), z0.device)
    hint.upstream.fill_(1.0)
    wts = (1.0, 0.1, 1.0, 0.5)
    z = z0.clone().requires_grad_(True)
    assert L.lean_ok(z, y, 15, False, hint)
    lean = L.fused_terms_split(z, y, t, region='ignored', terms=15, grad_hint=hint)
    assert all(v.dim() == 0 and v.dtype == torch.float32 for v in lean)
    assert type(lean[0].grad_fn).__name__.startswith('FusedTermsLean')
    sum(w * v for w, v in zip(wts, lean)).backward()
    z2 = z0.clone().requires_grad_(True)
    comp = L.fused_terms(z2, y, t, region='ignored', terms=15, grad_hint=hint)
    sum(w * comp[k] for k, w in enumerate(wts)).backward()
    assert [v.item() for v in lean] == [comp[k].item() for k in range(4)]
    assert torch.equal(z.grad, z2.grad)
    # only two of the four terms enter the loss
    z3 = z0.clone().requires_grad_(True)
    lean = L.fused_terms_split(z3, y, t, region='ignored', terms=15, grad_hint=hint)
    (lean[0] + 0.5 * lean[3]).backward()
    z4 = z0.clone().requires_grad_(True)
    ref = oloss.compute_loss(z4, y, t, w_kld=0.0, w_ent=0.0)          # CE + 0.5 SoftCE('ignored') only
    sum(ref.values()).backward()
    grad_close(z3.grad, z4.grad)
    # weighted outputs: w_k * term_k out of the kernels == the products composed by torch, values and gradient bits
    z5 = z0.clone().requires_grad_(True)
    wl = L.fused_terms_split(z5, y, t, region='ignored', terms=15, grad_hint=hint, weighted=True)
    z6 = z0.clone().requires_grad_(True)
    pl = L.fused_terms_split(z6, y, t, region='ignored', terms=15, grad_hint=hint)
    prods = [w * v for w, v in zip(wts, pl)]
    assert [v.item() for v in wl] == [v.item() for v in prods]
    (sum(wl) * 3.0).backward()
    (sum(prods) * 3.0).backward()
    assert torch.equal(z5.grad, z6.grad)
