"""Fused single-kernel IAS window (hiast_ias_fused_window) == the three-kernel pipeline, bit for bit.

Reference path: workflows/pseudo_label_generator.py:181-213 (the loop body of IASPseudoGenerator.run); the
three-kernel pipeline is itself pinned against the oracle / reference fixtures in test_ias_gpu.py."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import ias as oias

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _needs_development_build():
    """The fused persistent kernel measured 40 % slower than the three kernels and is compiled only
    into the development build (HIAST_DEV_VARIANTS=1 python -m hiast_b200.build --force)."""
    from hiast_b200 import _lib
    if not _lib.lib().hiast_dev_variants():
        pytest.skip('fused window kernel: development build only')


def engines(C, H, W, B, n, **kw):
    from hiast_b200.ias_engine import IASEngine
    n = ((n + B - 1) // B) * B
    fused = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, n, fused=True, **kw)
    split = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, n, fused=False, **kw)
    return fused, split


def assert_same(fused, split, n, B):
    g = (n + B - 1) // B
    assert torch.equal(fused.thr_groups[:g], split.thr_groups[:g])
    assert torch.equal(fused.temp_groups[:g], split.temp_groups[:g])
    assert torch.equal(fused.thr_state, split.thr_state)
    assert torch.equal(fused.plbl[:n], split.plbl[:n])
    assert torch.equal(fused.counts[:n], split.counts[:n])
    assert torch.equal(fused.confsum[:g], split.confsum[:g])
    assert torch.equal(fused.mean_state, split.mean_state)
    assert fused.check_errors() == split.check_errors()


@pytest.mark.parametrize('shape', [(5, 19, 64, 128, 2), (7, 16, 40, 64, 3), (1, 19, 8, 16, 2), (64, 19, 32, 64, 2),
                                   (9, 19, 96, 160, 4)])
def test_fused_equals_three_kernels_small(shape):
    n, C, H, W, B = shape
    g = torch.Generator().manual_seed(sum(shape))
    logits = torch.cat([gi.diffuse_logits(g, (n + 1) // 2, C, H, W), gi.peaked_logits(g, n // 2, C, H, W)] if n > 1
                       else [gi.peaked_logits(g, 1, C, H, W)]).cuda()
    fused, split = engines(C, H, W, B, n)
    for _ in range(3):                      # state carried over three windows
        fused.process(logits)
        split.process(logits)
        assert_same(fused, split, n, B)


def test_fused_against_oracle():
    """Same check as __graft_entry__.smoke(): the fused window against the CPU oracle on torch's CUDA softmax."""
    spec = gi.IAS_SPECS['ias_small']
    C, B = spec['C'], spec['B']
    batches = gi.ias_batches(spec)
    logits = torch.cat([lg for lg, _ in batches]).cuda()
    n, _, H, W = logits.shape
    from hiast_b200.ias_engine import IASEngine
    eng = IASEngine(C, H, W, B, spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'], ((n + B - 1) // B) * B)
    assert eng.process_fused(logits)
    eng.mean_prob(0, n)
    oracle = oias.IASOracle(C, spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    G = len(batches)
    assert np.array_equal(eng.thr_groups[:G].cpu().numpy(), np.stack(oracle.threshold_trace))
    assert np.array_equal(eng.plbl[:n].cpu().numpy(), np.stack(oracle.labels))
    assert np.array_equal(eng.counts[:n].sum(0).cpu().numpy(), oracle.statics_class)
    np.testing.assert_allclose(eng.mean_state.cpu().numpy(), oracle.class_mean_probs, rtol=1e-6)
    assert eng.check_errors()


@pytest.mark.parametrize('gif', [0, 1, 4])
def test_fused_full_resolution(gif):
    """19x1024x2048 (BASELINE.json configs[1] shape): 7 images (a trailing 1-image group), every groups-in-flight
    setting, with and without discarding the consumed spill lines."""
    from hiast_b200 import ops
    n, C, H, W, B = 7, 19, 1024, 2048, 2
    g = torch.Generator(device='cuda').manual_seed(77)
    logits = torch.randn(n, C, H, W, generator=g, device='cuda') * 3
    low = torch.randn(3, C, 32, 64, generator=g, device='cuda')
    logits[1] = torch.nn.functional.interpolate(low[:1] * 4, size=(H, W), mode='bilinear', align_corners=True)[0] + logits[1] / 6
    logits[2] = torch.nn.functional.interpolate(low[1:2] * 60, size=(H, W), mode='bilinear', align_corners=True)[0] + logits[2] / 6
    logits[5] = torch.linspace(-1.0, 1.0, C, device='cuda')[:, None, None]          # one bin takes a whole image
    fused, split = engines(C, H, W, B, 8)
    fused.groups_in_flight = gif
    split.process(logits)
    fused.process(logits)
    assert_same(fused, split, n, B)
    # size-independent properties of the fused result
    assert int((fused.counts[:n].sum())) == int((fused.plbl[:n] != 255).sum())
    # keep_spill: conf / label stay defined and equal phase A's
    ws = ops.ias_fused_workspace(n, B, 'cuda')
    thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
    e = fused
    e.counts.zero_(); e.confsum.zero_()
    assert ops.ias_fused_window(logits, B, e.key_lo, 0.5, 0.9, 8.0, e.conf[:n], e.label[:n], e.hist[:4], thr_state, e.thr_groups[:4],
                                e.temp_groups[:4], e.plbl[:n], e.counts[:n], e.confsum[:4], e.error_flag, ws, keep_spill=True)
    conf, label, _ = ops.ias_softmax_hist(logits, B)
    assert torch.equal(e.conf[:n], conf) and torch.equal(e.label[:n], label)
    assert torch.equal(e.plbl[:n], split.plbl[:n])


def test_fused_unsupported_shape_falls_back():
    n, C, H, W, B = 3, 7, 31, 51, 2           # C not in {16, 19}, HW % 4 != 0
    g = torch.Generator().manual_seed(5)
    logits = gi.diffuse_logits(g, n, C, H, W).cuda()
    fused, split = engines(C, H, W, B, 4)
    assert fused.process_fused(logits) is False
    fused.process(logits)                      # three-kernel fall-back inside
    split.process(logits)
    assert_same(fused, split, n, B)
