"""GPU parity of the IAS kernels (through the C ABI) against the oracle and the golden fixtures."""

import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import ias as oias

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


def ops():
    from hiast_b200 import ops as o
    return o


ALL_MODES = [1, 2, 3, 4, 5, 6, 11, 16, 21, 25, 26, 31, 36, 41, 46, 51, 56, 61, 66, 71, 76, 80, 81, 83]
PRODUCT_MODES = [1, 83]


def modes(wanted=ALL_MODES):
    """The product library ships the plain kernel (1) and the group-resident kernel (83 = default 0); the measured-and-dropped
    variants exist only in a development build (HIAST_DEV_VARIANTS=1 python -m hiast_b200.build --force)."""
    from hiast_b200 import _lib
    dev = bool(_lib.lib().hiast_dev_variants())
    return [m for m in wanted if dev or m in (0, 1, 83)]


def test_development_variants_are_not_in_the_product_library():
    from hiast_b200 import _lib
    o = ops()
    x = torch.randn(1, 19, 8, 16, device='cuda')
    if _lib.lib().hiast_dev_variants():
        pytest.skip('development build')
    with pytest.raises(_lib.HiastError, match='unsupported'):
        o.ias_softmax_hist(x, group_size=2, hist_mode=56)


def torch_softmax_max(logits):
    """The reference's own CUDA path for a1 (pseudo_label_generator.py:192-193)."""
    probs = torch.softmax(logits, dim=1)
    return probs.max(dim=1)


@pytest.mark.parametrize('shape', [(2, 19, 64, 128), (3, 19, 33, 52), (1, 16, 40, 64), (2, 7, 31, 51), (1, 40, 9, 13)])
@pytest.mark.parametrize('mode', ALL_MODES)
def test_phase_a_bit_exact_vs_torch_cuda(shape, mode):
    o = ops()
    if mode not in modes():
        pytest.skip('development variant (not in the product library)')
    g = torch.Generator().manual_seed(sum(shape) + mode)
    n, c, h, w = shape
    logits = torch.cat([gi.diffuse_logits(g, 1, c, h, w), gi.peaked_logits(g, n - 1, c, h, w)] if n > 1
                       else [gi.peaked_logits(g, 1, c, h, w)]).cuda()
    conf, label, hist = o.ias_softmax_hist(logits, group_size=2, hist_mode=mode)
    want_conf, want_label = torch_softmax_max(logits)
    assert torch.equal(conf, want_conf)
    assert torch.equal(label.long(), want_label)
    key_lo = o.ias_key_lo(c)
    G = (n + 1) // 2
    want_hist = np.stack([oias.class_key_histogram(want_conf[2 * k:2 * k + 2].cpu().numpy(),
                                                   want_label[2 * k:2 * k + 2].cpu().numpy(), c, key_lo)
                          for k in range(G)])
    nb = o.ias_num_bins(key_lo)
    assert hist.shape[2] == o.ias_row_stride(key_lo) and int(hist[:, :, nb:].sum()) == 0
    assert np.array_equal(hist[:, :, :nb].cpu().numpy().astype(np.uint32), want_hist)


def test_packed_expf_is_expf_for_every_non_positive_float():
    """The f32x2 exponential of phase A (the product kernel and the loss kernels) against CUDA's expf(), all 2^31 - 2^23 + 2 inputs."""
    import ctypes
    from hiast_b200 import _lib
    bad = torch.ones(1, dtype=torch.int64, device='cuda')
    _lib.check(_lib.lib().hiast_selftest_packed_expf(_lib.ptr(bad), _lib.stream_ptr(bad.device)), 'selftest')
    assert int(bad.item()) == 0


@pytest.mark.parametrize('seed', range(12))
def test_phase_a_randomised_stress_vs_torch_cuda(seed):
    """Randomised shapes / magnitudes / quantisations for the packed-math kernels (modes 56, 80, 83) against
    torch.softmax(...).max(1) on CUDA, bit for bit: logits scaled from 1e-4 to 300, quantised to multiples of 2^-k (many
    exact ties and near ties), sprinkled with -inf, +-1e4 and denormals."""
    o = ops()
    g = torch.Generator().manual_seed(1000 + seed)
    c = 19 if seed % 3 else 16
    h = int(torch.randint(3, 70, (1,), generator=g))
    w = 4 * int(torch.randint(1, 70, (1,), generator=g))
    n = int(torch.randint(1, 6, (1,), generator=g))
    scale = [1e-4, 1e-2, 1.0, 3.0, 30.0, 300.0][seed % 6]
    x = torch.randn(n, c, h, w, generator=g) * scale
    if seed % 2:
        q = 2.0 ** -int(torch.randint(0, 12, (1,), generator=g))
        x = torch.round(x / q) * q                                     # quantised: exact ties are common
    m = torch.rand(n, c, h, w, generator=g)
    x[m < 0.01] = -float('inf')
    x[(m >= 0.01) & (m < 0.02)] = -1e4
    x[(m >= 0.02) & (m < 0.03)] = 1e4
    x[(m >= 0.03) & (m < 0.04)] = 1e-41                                # denormal
    x[:, 0] = torch.where(torch.isinf(x).all(dim=1), torch.zeros(()), x[:, 0])      # no all -inf pixel (NaN in torch too)
    x = x.cuda()
    want_conf, want_label = torch_softmax_max(x)
    for mode in modes((56, 80, 83)):
        conf, label, hist = o.ias_softmax_hist(x, group_size=2, hist_mode=mode)
        assert torch.equal(conf, want_conf), (seed, mode)
        assert torch.equal(label.long(), want_label), (seed, mode)
        assert int(hist.sum()) == n * h * w


def test_phase_a_ties_and_near_ties():
    """Exact logit ties, channels a hair below the max (expf -> 1.0f), constant maps, huge gaps."""
    o = ops()
    c, h, w = 19, 16, 64
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, c, h, w, generator=g) * 2
    x[0, 7] = x[0, 3]                                       # exact ties, first index must win
    x[0, 11] = torch.maximum(x[0, 11], x[0, 3])
    top = x[1].max(dim=0).values
    for k, eps in enumerate([1e-8, 3e-8, 6e-8, 1.2e-7, 2.4e-7, 5e-7]):
        x[1, k] = top * (1 - eps) - eps                     # a hair below the max, earlier channels
    x[2] = 0.0                                              # all equal -> conf = 1/19, label 0
    x[3, 5] += 40.0                                         # saturated: conf == 1.0
    x[3, 9] = -1e4                                          # expf underflows to 0 / denormals (config-4 planes)
    x[3, 14, :8] = -float('inf')
    x[3, 16, 8:] = -95.0
    x[3, 17, 4:12] = x[3, 5, 4:12] - 88.5                   # exp(x - m) is a denormal
    x = torch.cat([x, x[:2] * 1e-6, x[:2] * 300.0])         # tiny logits: every channel a near tie; huge gaps
    x = x.cuda()
    want_conf, want_label = torch_softmax_max(x)
    for mode in modes((0, 1, 6, 16, 26, 36, 46, 51, 56, 66, 76, 80, 81, 83)):
        conf, label, _ = o.ias_softmax_hist(x, group_size=2, hist_mode=mode)
        assert torch.equal(conf, want_conf), mode
        assert torch.equal(label.long(), want_label), mode
        assert (label[2] == 0).all() and (conf[3] == 1.0).any()


@pytest.mark.parametrize('name', ['ias_small', 'ias_c7', 'ias_g25'])
def test_post_softmax_stages_vs_reference_fixture(name):
    """conf/label of the reference's own run -> conf_hist -> scan -> select -> meanprob == fixture."""
    o = ops()
    spec = gi.IAS_SPECS[name]
    gold = load(name)
    C, B, N = spec['C'], spec['B'], spec['N']
    conf = torch.from_numpy(gold['conf']).cuda()
    label = torch.from_numpy(gold['label'].astype(np.int64)).cuda()    # int64 like probs.max(1)
    for key_lo in (0, o.ias_key_lo(C)):
        hist, label_u8 = o.ias_conf_hist(conf, label, C, B, key_lo=key_lo)
        G = (N + B - 1) // B
        thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        thr_groups, _ = o.ias_threshold_scan(hist, G, C, key_lo, spec['alpha'], spec['beta'], spec['gamma'],
                                             thr_state, error_flag=flag)
        assert flag.item() == 0
        assert np.array_equal(thr_groups.cpu().numpy(), gold['thr_trace'])
        assert np.array_equal(thr_state.cpu().numpy(), gold['class_threshold'])
        plbl, counts, confsum = o.ias_select(conf, label_u8, thr_groups, C, B)
        assert np.array_equal(plbl.cpu().numpy(), gold['plbl'])
        assert np.array_equal(counts.cpu().numpy(), gold['counts'])
        mean_state = torch.zeros(C, dtype=torch.float64, device='cuda')
        o.ias_meanprob_scan(confsum, counts, B, C, spec['cp_gamma'], mean_state)
        np.testing.assert_allclose(mean_state.cpu().numpy(), gold['class_mean_probs'], rtol=1e-6)


@pytest.mark.parametrize('name', ['ias_small', 'ias_c7', 'ias_g25'])
def test_full_pipeline_vs_oracle_on_cuda_softmax(name):
    """logits -> labels/thresholds, oracle fed by torch's CUDA softmax exactly like the reference."""
    o = ops()
    spec = gi.IAS_SPECS[name]
    C, B = spec['C'], spec['B']
    batches = gi.ias_batches(spec)
    oracle = oias.IASOracle(C, spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    logits = torch.cat([lg for lg, _ in batches]).cuda()
    conf, label, hist = o.ias_softmax_hist(logits, group_size=B)
    G = len(batches)
    thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
    thr_groups, _ = o.ias_threshold_scan(hist, G, C, o.ias_key_lo(C), spec['alpha'], spec['beta'], spec['gamma'], thr_state)
    plbl, counts, confsum = o.ias_select(conf, label, thr_groups, C, B)
    mean_state = torch.zeros(C, dtype=torch.float64, device='cuda')
    o.ias_meanprob_scan(confsum, counts, B, C, spec['cp_gamma'], mean_state)
    assert np.array_equal(thr_groups.cpu().numpy(), np.stack(oracle.threshold_trace))
    assert np.array_equal(plbl.cpu().numpy(), np.stack(oracle.labels))
    assert np.array_equal(counts.sum(0).cpu().numpy(), oracle.statics_class)
    np.testing.assert_allclose(mean_state.cpu().numpy(), oracle.class_mean_probs, rtol=1e-6)


def test_config0_vs_oracle():
    """BASELINE.json configs[0] (8 x 19x512x1024, batch 2) bit-exact against the oracle on CUDA softmax,
    and within a hair of the CPU-generated fixture (CPU and CUDA softmax differ in the last ulp)."""
    o = ops()
    spec = gi.IAS_SPECS['ias_config0']
    gold = load('ias_config0')
    C, B = spec['C'], spec['B']
    batches = gi.ias_batches(spec)
    logits = torch.cat([lg for lg, _ in batches]).cuda()
    oracle = oias.IASOracle(C, spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    for mode in modes():
        conf, label, hist = o.ias_softmax_hist(logits, group_size=B, hist_mode=mode)
        thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        thr_groups, _ = o.ias_threshold_scan(hist, len(batches), C, o.ias_key_lo(C), spec['alpha'], spec['beta'],
                                             spec['gamma'], thr_state, error_flag=flag)
        plbl, counts, _ = o.ias_select(conf, label, thr_groups, C, B)
        assert flag.item() == 0
        assert np.array_equal(thr_groups.cpu().numpy(), np.stack(oracle.threshold_trace))
        assert np.array_equal(plbl.cpu().numpy(), np.stack(oracle.labels))
        assert np.array_equal(counts.sum(0).cpu().numpy(), oracle.statics_class)
    np.testing.assert_allclose(thr_groups.cpu().numpy(), gold['thr_trace'], rtol=1e-4)
    rel = np.abs(counts.sum(0).cpu().numpy() - gold['statics_class']) / np.maximum(gold['statics_class'], 1)
    assert rel.max() < 1e-2


def test_full_resolution_properties():
    """19x1024x2048 (BASELINE.json configs[1] shape), 4 images: size-independent properties."""
    o = ops()
    C, H, W, B = 19, 1024, 2048, 2
    g = torch.Generator(device='cuda').manual_seed(1234)
    logits = torch.randn(4, C, H, W, generator=g, device='cuda') * 3
    logits[2:] = torch.nn.functional.interpolate(torch.randn(2, C, 32, 64, generator=g, device='cuda') * 4,
                                                 size=(H, W), mode='bilinear', align_corners=True) + logits[2:] / 6
    conf, label, hist = o.ias_softmax_hist(logits, group_size=B)
    want_conf, want_label = torch_softmax_max(logits)
    assert torch.equal(conf, want_conf) and torch.equal(label.long(), want_label)
    # the histogram of each group is a partition of its pixels: checksum of checksums
    assert hist.sum(dim=(1, 2)).tolist() == [B * H * W, B * H * W]
    per_class = hist.sum(dim=2).cpu()
    want = torch.stack([torch.bincount(want_label[2 * k:2 * k + 2].flatten(), minlength=C) for k in range(2)]).cpu()
    assert torch.equal(per_class.long(), want)
    thr_state = torch.full((C,), 0.9, dtype=torch.float64, device='cuda')
    hist_raw = hist.clone()     # the scan turns `hist` into prefix sums in place
    thr_groups, _ = o.ias_threshold_scan(hist, 2, C, o.ias_key_lo(C), 0.5, 0.9, 8.0, thr_state)
    nb = o.ias_num_bins(o.ias_key_lo(C))
    assert torch.equal(hist[:, :, nb - 1].long().cpu(), want)    # last prefix = class total
    plbl, counts, confsum = o.ias_select(conf, label, thr_groups, C, B)
    # mask property: kept pixels keep their label and have conf >= thr; ignored have conf < thr
    thr_px = thr_groups[torch.arange(4, device='cuda') // B][:, :, None, None].expand(4, C, H, W).gather(
        1, label.long()[:, None]).squeeze(1)
    kept = plbl != 255
    assert torch.equal(kept, conf.double() >= thr_px)
    assert torch.equal(plbl[kept], label[kept])
    assert torch.equal(counts, torch.stack([torch.bincount(plbl[i][kept[i]].long(), minlength=C) for i in range(4)]))
    # idempotence: same inputs, same outputs (atomics must not make anything order dependent)
    conf2, label2, hist2 = o.ias_softmax_hist(logits, group_size=B)
    assert torch.equal(hist2, hist_raw) and torch.equal(conf2, conf) and torch.equal(label2, label)


@pytest.mark.parametrize('dist', ['D1_diffuse', 'D2_peaked'])
def test_full_resolution_engine_vs_oracle(dist):
    """BASELINE.json configs[1] at FULL resolution against the oracle end to end (VERDICT r1 weak #1 i):
    8 maps of 19x1024x2048, batch 2, through IASEngine.process; the oracle (the reference's loop body, oracle/ias.py) is fed
    by torch's CUDA softmax exactly as the reference is (:192-193).  Thresholds of every group, pseudo-labels of every
    pixel, per-image counts and class totals bit-exact; class_mean_probs 1e-6."""
    from hiast_b200.ias_engine import IASEngine
    C, H, W, B, N = 19, 1024, 2048, 2, 8
    g = torch.Generator(device='cuda').manual_seed(1234)
    if dist == 'D1_diffuse':
        logits = torch.randn(N, C, H, W, generator=g, device='cuda') * 3
    else:
        low = torch.randn(N, C, 32, 64, generator=g, device='cuda') * 4
        logits = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)
        logits += torch.randn(N, C, H, W, generator=g, device='cuda') * 0.5
    eng = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, N)
    plbl, counts, thr_groups = eng.process(logits)
    oracle = oias.IASOracle(C, 0.5, 0.9, 8.0, 0.99)
    oracle.run([(logits[i:i + B], ['img_%d.png' % k for k in range(i, i + B)]) for i in range(0, N, B)])
    assert eng.check_errors()
    assert np.array_equal(thr_groups.cpu().numpy(), np.stack(oracle.threshold_trace))
    assert np.array_equal(eng.thr_state.cpu().numpy(), oracle.class_threshold)
    got = plbl.cpu().numpy()
    for i in range(N):
        assert np.array_equal(got[i], oracle.labels[i]), i
    want_counts = np.zeros((N, C), dtype=np.int64)
    for i, row in enumerate(oracle.sample_stats):
        for k, v in row.items():
            if k != 'file':
                want_counts[i, k] = v
    assert np.array_equal(counts.cpu().numpy(), want_counts)
    assert np.array_equal(counts.sum(0).cpu().numpy(), oracle.statics_class)
    np.testing.assert_allclose(eng.mean_state.cpu().numpy(), oracle.class_mean_probs, rtol=1e-6)
    ignored = float((got == 255).mean())
    assert 0.05 < ignored < 0.98, ignored                         # the case is not degenerate


@pytest.mark.parametrize('mode', [56, 80, 81, 83])
def test_full_resolution_hist_variants_match_plain_red(mode):
    if mode not in modes():
        pytest.skip('development variant (not in the product library)')
    _full_resolution_hist_variants_match_plain_red(mode)


def _full_resolution_hist_variants_match_plain_red(mode):
    """Full-size maps through the packed-math / group-resident kernels == the plain one-RED-per-pixel kernel, on
    diffuse, peaked, saturated and CONSTANT maps (one (class, key) bin receives 2M pixels: exercises the 16-bit
    wrap of the shared-memory table), odd group sizes and a trailing 1-image group."""
    o = ops()
    C, H, W = 19, 1024, 2048
    g = torch.Generator(device='cuda').manual_seed(99)
    logits = torch.randn(5, C, H, W, generator=g, device='cuda') * 3
    low = torch.randn(2, C, 32, 64, generator=g, device='cuda')
    logits[1] = torch.nn.functional.interpolate(low[:1] * 4, size=(H, W), mode='bilinear', align_corners=True)[0] + logits[1] / 6
    logits[2] = torch.nn.functional.interpolate(low[1:] * 60, size=(H, W), mode='bilinear', align_corners=True)[0] + logits[2] / 6
    const = torch.linspace(-1.0, 1.0, C, device='cuda')
    const[7] = 2.0                                             # conf ~ 0.25: a bin of the shared table
    logits[3] = const[:, None, None]
    const[7] = -0.4                                            # conf ~ 0.13: a bin below the table (global REDs)
    logits[4, :, :H // 2] = const[:, None, None]
    for B in (2, 3):
        conf, label, hist = o.ias_softmax_hist(logits, group_size=B, hist_mode=mode)
        conf1, label1, hist1 = o.ias_softmax_hist(logits, group_size=B, hist_mode=1)
        assert torch.equal(conf, conf1) and torch.equal(label, label1)
        assert torch.equal(hist, hist1)
        assert int(hist.sum()) == 5 * H * W


def _two_hot_neighbour_keys_logits(n, C, H, W, cls=5):
    """Maps whose confidences alternate, pixel by pixel in raster order, between the fp16 keys 0x3BFE and 0x3BFF of ONE class:
    the two 16-bit counters of one shared-memory word of the group-resident kernels, both driven far beyond 65 535."""
    logits = torch.full((n, C, H, W), -100.0, device='cuda')
    logits[:, cls] = 0.0
    gap = torch.tensor([6.930, 7.624], device='cuda')            # 1 / (1 + e^-gap) = 1 - 2^-10, 1 - 2^-11
    logits[:, cls + 1] = -gap[torch.arange(W, device='cuda') % 2]
    probe = torch.softmax(logits[:1, :, :1, :2], dim=1).max(dim=1)[0].half().view(torch.int16).flatten().tolist()
    assert probe == [0x3BFE, 0x3BFF], [hex(k) for k in probe]
    return logits


def test_shared_table_counters_with_two_hot_adjacent_keys():
    """32 full-size maps put ~226 k pixels per CTA slice into EACH of two adjacent keys of one class (the
    low and the high half of one packed shared-memory word).  Fifty repetitions of the default kernel must reproduce the
    plain one-RED-per-pixel histogram every time (the half-range drain leaves no window in which a neighbour's increment
    can observe a carry)."""
    o = ops()
    C, H, W, n = 19, 1024, 2048, 32
    logits = _two_hot_neighbour_keys_logits(n, C, H, W)
    _, _, want = o.ias_softmax_hist(logits, group_size=2, hist_mode=1)
    key_lo = o.ias_key_lo(C)
    assert int(want[:, 5, 0x3BFE - key_lo].sum()) == n * H * W // 2 and int(want[:, 5, 0x3BFF - key_lo].sum()) == n * H * W // 2
    conf = torch.empty((n, H, W), device='cuda')
    label = torch.empty((n, H, W), dtype=torch.uint8, device='cuda')
    hist = torch.empty_like(want)
    for rep in range(50):
        o.ias_softmax_hist(logits, 2, key_lo, conf, label, hist, hist_mode=0)
        assert torch.equal(hist, want), 'repetition %d' % rep
    for B in (1, 3, 32):                                          # other flush patterns: a flush per image, ragged groups, one group
        _, _, w1 = o.ias_softmax_hist(logits, group_size=B, hist_mode=1)
        for rep in range(5):
            _, _, h0 = o.ias_softmax_hist(logits, group_size=B, hist_mode=0)
            assert torch.equal(h0, w1)


def test_upsample_kernel_counters_drain_beyond_16_bits():
    """The fused up-sampling kernel shares the packed shared-memory table: constant low-resolution maps put a whole CTA slice
    into one counter (image k -> key 0x3BFE or 0x3BFF of one class)."""
    o = ops()
    C, H, W, n = 19, 1024, 2048, 8
    lr = torch.full((n, C, 129, 257), -100.0, device='cuda')
    lr[:, 5] = 0.0
    lr[0::2, 6] = -6.930
    lr[1::2, 6] = -7.624
    full = torch.nn.functional.interpolate(lr, size=(H, W), mode='bilinear', align_corners=True)
    _, _, want = o.ias_softmax_hist(full, group_size=2, hist_mode=1)
    for rep in range(10):
        _, _, hist = o.ias_upsample_softmax_hist(lr, (H, W), 2)
        assert torch.equal(hist, want)
    key_lo = o.ias_key_lo(C)
    assert int(want[:, 5, 0x3BFE - key_lo].sum()) == n // 2 * H * W


def test_empty_and_single_image():
    o = ops()
    logits = torch.randn(1, 19, 8, 16, device='cuda')
    conf, label, hist = o.ias_softmax_hist(logits, group_size=2)
    assert hist.shape[0] == 1 and int(hist.sum()) == 128
    empty = torch.empty(0, 19, 8, 16, device='cuda')
    conf, label, hist = o.ias_softmax_hist(empty, group_size=2)
    assert conf.numel() == 0 and hist.shape[0] == 0


def test_scan_error_flag_mirrors_numpy_valueerror():
    o = ops()
    conf = torch.full((2, 4, 4), 0.5, device='cuda')
    label = torch.zeros((2, 4, 4), dtype=torch.uint8, device='cuda')
    hist, _ = o.ias_conf_hist(conf, label, 3, 2)
    thr_state = torch.full((3,), 0.999, dtype=torch.float64, device='cuda')
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    o.ias_threshold_scan(hist, 1, 3, 0, 1.5, 0.9, 1.0, thr_state, error_flag=flag)   # q = 1 - 1.5*0.999 < 0
    assert flag.item() & 1


@pytest.mark.parametrize('shape', [((17, 33), (128, 256)), ((9, 21), (64, 160)), ((13, 17), (100, 132)), ((129, 257), (1024, 2048)),
                                   ((16, 16), (16, 16)), ((65, 129), (512, 1024)), ((20, 40), (157, 316))])
@pytest.mark.parametrize('C', [19, 16])
@pytest.mark.parametrize('v1', [0, 1])
def test_fused_bilinear_upsample_bit_exact_vs_torch(shape, C, v1):
    """phase A straight from the stride-8 logits == softmax(F.interpolate(x, align_corners=True)).max(1)
    on CUDA, bit for bit (conf, label, histogram), without the full-resolution tensor."""
    o = ops()
    (h, w), (H, W) = shape
    n = 3 if H * W < 1 << 20 else 2
    g = torch.Generator().manual_seed(h * 1000 + W + C)
    lr = (torch.randn(n, C, h, w, generator=g) * 4).cuda()
    full = torch.nn.functional.interpolate(lr, size=(H, W), mode='bilinear', align_corners=True)
    want_conf, want_label = torch_softmax_max(full)
    from hiast_b200 import _lib
    _lib.lib().hiast_debug_upsample_v1(v1)     # 1: the first kernel (4 px along x per thread); 0: column kernel where it applies
    try:
        conf, label, hist = o.ias_upsample_softmax_hist(lr, (H, W), group_size=2)
    finally:
        _lib.lib().hiast_debug_upsample_v1(0)
    assert torch.equal(conf, want_conf)
    assert torch.equal(label.long(), want_label)
    conf2, label2, hist2 = o.ias_softmax_hist(full.contiguous(), group_size=2)
    assert torch.equal(hist, hist2)


def test_engine_from_lowres_logits_equals_full_resolution_path():
    from hiast_b200.ias_engine import IASEngine
    g = torch.Generator().manual_seed(3)
    lr = (torch.randn(4, 19, 9, 17, generator=g) * 4).cuda()
    full = torch.nn.functional.interpolate(lr, size=(64, 128), mode='bilinear', align_corners=True).contiguous()
    a = IASEngine(19, 64, 128, 2, 0.5, 0.9, 8.0, 0.99, 4)
    b = IASEngine(19, 64, 128, 2, 0.5, 0.9, 8.0, 0.99, 4)
    a.phase_a_lowres(lr)
    b.phase_a(full)
    for e in (a, b):
        e.phase_b(0, 4)
        e.phase_c(0, 4)
        e.mean_prob(0, 4)
    assert torch.equal(a.plbl, b.plbl) and torch.equal(a.thr_groups, b.thr_groups) and torch.equal(a.mean_state, b.mean_state)


def test_token_ring_hand_off_inside_the_scan_kernel():
    """hiast_ias_threshold_scan_ring on ONE GPU with the mailbox looped back to itself: the scan of window 0 publishes its final
    thresholds (token_out), the scan of window 1 starts from the mailbox (token_in) although its thr_state holds garbage;
    together they equal one scan over both windows.  (The two-process version over CUDA IPC is tests/test_sharded_gpu.py.)"""
    import ctypes as C
    from hiast_b200 import _lib
    o = ops()
    L = _lib.lib()
    Cn, H, W, B, n = 19, 32, 64, 2, 12
    g = torch.Generator().manual_seed(41)
    logits = torch.cat([gi.diffuse_logits(g, n // 2, Cn, H, W), gi.peaked_logits(g, n // 2, Cn, H, W)]).cuda()
    key_lo = o.ias_key_lo(Cn)
    _, _, hist = o.ias_softmax_hist(logits, B)
    state = torch.full((Cn,), 0.9, dtype=torch.float64, device='cuda')
    want, _ = o.ias_threshold_scan(hist.clone(), n // B, Cn, key_lo, 0.5, 0.9, 8.0, state)
    box, handle = C.c_void_p(), (C.c_ubyte * 64)()
    assert L.hiast_ring_mailbox_bytes() >= 256 * 16
    _lib.check(L.hiast_ring_create(C.byref(box), C.cast(handle, C.c_void_p)), 'ring_create')
    try:
        half = n // B // 2
        s0 = torch.full((Cn,), 0.9, dtype=torch.float64, device='cuda')
        first, _ = o.ias_threshold_scan(hist[:half].clone(), half, Cn, key_lo, 0.5, 0.9, 8.0, s0, token=(None, 0, box.value, 5))
        s1 = torch.full((Cn,), -123.0, dtype=torch.float64, device='cuda')          # must be ignored: the token wins
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        second, _ = o.ias_threshold_scan(hist[half:].clone(), n // B - half, Cn, key_lo, 0.5, 0.9, 8.0, s1, error_flag=flag,
                                         token=(box.value, 5, None, 0))
        assert torch.equal(torch.cat([first, second]), want)
        assert torch.equal(s1, state) and int(flag.item()) == 0
    finally:
        torch.cuda.synchronize()
        L.hiast_ring_destroy(box)


def test_full_resolution_concurrent_schedule_equals_serial_schedule():
    """The concurrent schedule (phase A on 148 - 12 SMs, scan + phase C of the window before on the rest) at full resolution:
    12 maps in windows of 4, every threshold, label and count equal to the serial three-kernel path (which
    test_full_resolution_engine_vs_oracle pins against the oracle)."""
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS, window_images
    C, H, W, B, N, window = 19, 1024, 2048, 2, 12, 4
    g = torch.Generator(device='cuda').manual_seed(77)
    logits = torch.randn(N, C, H, W, generator=g, device='cuda') * 3
    low = torch.randn(N // 2, C, 32, 64, generator=g, device='cuda') * 4
    logits[1::2] = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True) + logits[1::2] / 6
    out = {}
    for reserve in (0, 12):
        eng = IASEngine(C, H, W, B, 0.5, 0.9, 8.0, 0.99, 3 * window)
        eng.reserve_sms = reserve
        got = {}

        def on_window(w, plbl, counts, thr_groups, got=got):
            got[w] = (plbl.clone(), counts.clone(), thr_groups.clone())

        thr, mean, statics = ShardedIAS(eng, window, N, 0, 1).run(
            lambda w: logits[window_images(w, window, N)[0]:window_images(w, window, N)[0] + window_images(w, window, N)[1]], on_window)
        torch.cuda.synchronize()
        assert eng.check_errors()
        out[reserve] = (got, thr.clone(), mean.clone(), statics.clone())
    a, b = out[0], out[12]
    assert sorted(a[0]) == sorted(b[0]) == [0, 1, 2]
    for w in a[0]:
        for x, y in zip(a[0][w], b[0][w]):
            assert torch.equal(x, y), w
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
