"""Pin the oracle: every oracle function against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""

import ast
import hashlib
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import copy_paste as ocp
from oracle import ias as oias
from oracle import losses as oloss
from oracle import metrics as omet

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=False)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_inputs(batches, gold):
    got = [sha(lg.numpy()) for lg, _ in batches]
    if got != list(gold['logits_sha']):
        pytest.skip('seeded torch CPU generator does not reproduce the fixture inputs on this host')


@pytest.mark.parametrize('name', ['ias_small', 'ias_c7', 'ias_g25'])
@pytest.mark.parametrize('faithful', [False, True])
def test_ias_oracle_matches_reference(name, faithful):
    spec = gi.IAS_SPECS[name]
    gold = load(name)
    assert ast.literal_eval(str(gold['spec'])) == spec
    batches = gi.ias_batches(spec)
    check_inputs(batches, gold)
    o = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'], faithful=faithful)
    o.run(batches)
    assert np.array_equal(np.stack(o.threshold_trace), gold['thr_trace'])
    assert np.array_equal(o.class_threshold, gold['class_threshold'])
    assert np.array_equal(o.class_mean_probs, gold['class_mean_probs'])
    assert np.array_equal(o.statics_class, gold['statics_class'])
    assert np.array_equal(np.stack(o.labels), gold['plbl'])
    counts = np.zeros_like(gold['counts'])
    for i, st in enumerate(o.sample_stats):
        for k, v in st.items():
            if k != 'file':
                counts[i, k] = v
    assert np.array_equal(counts, gold['counts'])


def test_ias_oracle_from_stored_conf():
    """The post-softmax stages alone, fed the reference's own conf/label arrays."""
    for name in ['ias_small', 'ias_c7', 'ias_g25']:
        spec = gi.IAS_SPECS[name]
        gold = load(name)
        o = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
        i = 0
        while i < spec['N']:
            b = min(spec['B'], spec['N'] - i)
            o.step_conf(gold['conf'][i:i + b], gold['label'][i:i + b].astype(np.int64), ['x'] * b)
            i += b
        assert np.array_equal(np.stack(o.threshold_trace), gold['thr_trace'])
        assert np.array_equal(np.stack(o.labels), gold['plbl'])
        assert np.array_equal(o.class_mean_probs, gold['class_mean_probs'])


def test_ias_oracle_config0():
    """BASELINE.json configs[0]: 8 x 19x512x1024, batch 2 (thresholds + label digests)."""
    spec = gi.IAS_SPECS['ias_config0']
    gold = load('ias_config0')
    batches = gi.ias_batches(spec)
    check_inputs(batches, gold)
    o = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    o.run(batches)
    assert np.array_equal(np.stack(o.threshold_trace), gold['thr_trace'])
    assert np.array_equal(o.class_mean_probs, gold['class_mean_probs'])
    assert np.array_equal(o.statics_class, gold['statics_class'])
    assert [sha(p) for p in o.labels] == list(gold['plbl_sha'])


def test_hist_form_equals_np_quantile():
    """The histogram restatement (what the CUDA scan implements) == np.quantile, bit for bit."""
    rs = np.random.RandomState(0)
    n_bad = 0
    for trial in range(400):
        m = int(rs.choice([0, 1, 2, 3, 17, 500, 5000]))
        mode = trial % 3
        if mode == 0:
            conf = rs.uniform(1 / 19, 1, size=m).astype(np.float32)
        elif mode == 1:
            conf = (1 - rs.exponential(0.002, size=m)).clip(0.06, 1).astype(np.float32)
        else:
            conf = rs.choice(np.array([0.25, 0.5, 0.9, 0.9004, 1.0], dtype=np.float32), size=m)
        thr = float(rs.choice([0.9, rs.uniform(0.05, 0.9999), 0.999, float(np.float16(0.9004))]))
        alpha = float(rs.choice([0.2, 0.5, 1.0]))
        gamma = float(rs.choice([1.0, 8.0, 2.5]))
        label = np.zeros(m, dtype=np.int64)
        want = oias.ias_quantile_thresholds(conf, label, np.array([thr]), 1, alpha, gamma)[0]
        for key_lo in (0, 0x2ABD):
            hist = oias.class_key_histogram(conf, label, 1, key_lo)[0]
            got = oias.threshold_from_hist(hist, key_lo, np.float64(thr), alpha, gamma)
            n_bad += int(got.tobytes() != want.tobytes())
    assert n_bad == 0


@pytest.mark.parametrize('name', list(gi.LOSS_SPECS))
def test_loss_oracle_matches_reference(name):
    spec = gi.LOSS_SPECS[name]
    gold = load(name)
    z = torch.from_numpy(gold['z']).requires_grad_(True)
    t = torch.from_numpy(gold['t'])
    plbl = torch.from_numpy(gold['plbl'])
    s_z = s_lbl = None
    if spec['source']:
        s_z = torch.from_numpy(gold['s_z']).requires_grad_(True)
        s_lbl = torch.from_numpy(gold['s_lbl'])
    out = oloss.compute_loss(z, plbl, t, s_z, s_lbl, w_seg=spec['w_seg'], w_kld=spec['w_kld'],
                             w_ent=spec['w_ent'], w_cst=spec['w_cst'], cst_region=spec['region'])
    assert list(out.keys()) == list(gold['keys'])
    vals = np.array([v.item() for v in out.values()])
    np.testing.assert_allclose(vals, gold['values'], rtol=1e-6)
    sum(v.mean() for v in out.values()).backward()
    np.testing.assert_allclose(z.grad.numpy(), gold['grad'], rtol=1e-5, atol=1e-10)
    if spec['source']:
        np.testing.assert_allclose(s_z.grad.numpy(), gold['s_grad'], rtol=1e-5, atol=1e-10)


@pytest.mark.parametrize('name', list(gi.METRIC_SPECS))
def test_metric_oracle_matches_reference(name):
    spec = gi.METRIC_SPECS[name]
    gold = load(name)
    inter, union, pred_after = omet.intersection_and_union(gold['pred'], gold['target'], spec['K'])
    assert np.array_equal(inter, gold['intersection']) and inter.dtype == np.float32
    assert np.array_equal(union, gold['union'])
    assert np.array_equal(pred_after.reshape(gold['pred'].shape), gold['pred_after'])
    cm = omet.confusion_matrix(gold['pred'], gold['target'], spec['K'])
    K = spec['K']
    assert np.array_equal(np.diag(cm)[:K].astype(np.float32), gold['intersection'])
    area_union = cm[:K, :].sum(1) + cm[:, :K].sum(0) - np.diag(cm)[:K]
    assert np.array_equal(area_union.astype(np.float32), gold['union'])


def test_copy_paste_oracle_matches_reference():
    spec = gi.COPY_PASTE_SPEC
    gold = load('copy_paste')
    ds = gi.CopyPasteDataset(spec)
    cv, hard = ocp.hard_classes(gi.copy_paste_class_value(spec), spec['selected'])
    probs = ocp.class_probs(cv)
    assert np.array_equal(hard, gold['hard'])
    assert np.array_equal(probs, gold['probs'])
    np.random.seed(spec['seed'])
    for i in range(spec['n_run']):
        img, lbl, _ = ds.load_data(i)

        def load_donor(name):
            d_img, d_lbl, _ = ds.load_data(ds.get_file_to_idx(name))
            return d_img, d_lbl

        o_img, o_lbl, o_mask, _ = ocp.run_original(img, lbl, hard, probs, ds.get_samples_with_class(),
                                                   load_donor, spec['C'])
        assert np.array_equal(o_img, gold['img_%d' % i])
        assert np.array_equal(o_lbl, gold['lbl_%d' % i])
        assert np.array_equal(o_mask, gold['mask_%d' % i])


def test_cst_variant_oracles_match_reference():
    """LOSS['KLDIV'] / LOSS['MSE'] restatements against the reference's own values and gradients."""
    gold = load('loss_cst_variants')
    z0, t, tz, plbl = (torch.from_numpy(gold[k]) for k in ('z', 't', 'tz', 'plbl'))
    for kind, fn, tgt in (('kldiv', oloss.kl_div, tz), ('mse', oloss.mse, t)):
        for region in ('none', 'ignored', 'confident', 'all'):
            z = z0.clone().requires_grad_(True)
            val = fn(z, tgt) if region == 'none' else fn(z, tgt, refer_labels=plbl, region=region)
            val.backward()
            np.testing.assert_allclose(val.item(), gold['%s_%s' % (kind, region)], rtol=1e-6)
            np.testing.assert_allclose(z.grad.numpy(), gold['%s_%s_grad' % (kind, region)], rtol=1e-5, atol=1e-10)


def test_cbst_oracle_matches_reference():
    gold = load('cbst_small')
    spec = gi.IAS_SPECS['ias_small']
    batches = gi.ias_batches(spec)
    cl = [oias.softmax_max(lg) for lg, _ in batches]
    # installed-version semantic (numpy >= 2 evaluates this quantile in float16): bit-exact against the fixture
    thr = oias.cbst_thresholds(cl, spec['C'], int(gold['interval']), float(gold['p']), f64_quantile=False)
    assert np.array_equal(thr, gold['class_threshold'])
    labels = [oias.select_confident(c[k], l[k], thr).astype(np.uint8) for c, l in cl for k in range(len(c))]
    assert np.array_equal(np.stack(labels), gold['plbl'])
    # pinned-version semantic (float64 quantile, numpy 1.19): within one fp16 step of it on this small fixture
    thr64 = oias.cbst_thresholds(cl, spec['C'], int(gold['interval']), float(gold['p']))
    assert np.abs(thr64 - thr).max() < 2e-3


def test_ema_oracle_vs_reference_fixture():
    """oracle.ema == utils.update_ema_model (utils/utils.py:115-123) run by tests/golden/make_golden.py, bit for bit."""
    from oracle import ema
    g = np.load(os.path.join(GOLD, 'ema_update.npz'))
    n, nb = int(g['n_params']), int(g['n_buffers'])
    new = ema.ema_update([g['k%d' % i] for i in range(n)], [g['q%d' % i] for i in range(n)], float(g['gamma']))
    for i in range(n):
        assert np.array_equal(new[i], g['new%d' % i])
    for i, b in enumerate(ema.copy_buffers([g['bq%d' % i] for i in range(nb)])):
        assert np.array_equal(b, g['bnew%d' % i])


def test_ce_general_oracle_equals_reference_fixture():
    """LOSS['CE'] with class weights / refer_labels (losses.py:32-36,68-89): values and gradients, CPU torch both sides."""
    import golden_inputs as gi
    from oracle import losses as oloss
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'loss_ce_general.npz'))
    for key, spec in gi.CE_GENERAL_SPECS.items():
        z, labels, weights, refer = gi.ce_general_inputs(spec)
        for case, kw in gi.ce_general_cases(labels, weights, refer).items():
            zz = z.clone().requires_grad_(True)
            val = oloss.ce_general(zz, **kw)
            val.backward()
            assert val.item() == gold['%s_%s' % (key, case)], (key, case)
            assert np.array_equal(zz.grad.numpy(), gold['%s_%s_grad' % (key, case)]), (key, case)


def test_copy_paste_class_probs_and_draws_on_20_seeds():
    """the sampling probabilities (torch float64 arithmetic, preprocessor.py:29-34) and the class
    draws of random_select (:70-77) of the unmodified reference on 20 random class-value vectors, bit for bit -- for the
    oracle and for the product's host logic (hiast_b200.preprocessor.CopyPaste, no GPU involved)."""
    from types import SimpleNamespace
    from hiast_b200.preprocessor import CopyPaste, DonorSampler
    gold = np.load(os.path.join(GOLD, 'copy_paste_probs.npz'))
    for seed in range(int(gold['n_seeds'])):
        value, want_p, want_picks, hard = (gold['%s_%d' % (k, seed)] for k in ('value', 'probs', 'picks', 'hard'))
        C = len(value)
        assert np.array_equal(ocp.class_probs(value), want_p)
        cp = CopyPaste.__new__(CopyPaste)
        cp.cfg = SimpleNamespace(dataset=SimpleNamespace(num_classes=C))
        cp.class_value = value.copy()
        cp.class_probs = cp.calculate_class_probs()
        assert np.array_equal(cp.class_probs, want_p), seed
        np.random.seed(seed)
        assert [int(cp.random_select(hard)) for _ in range(len(want_picks))] == want_picks.tolist()
        np.random.seed(seed)
        assert [int(ocp.random_select(C, want_p, hard)) for _ in range(len(want_picks))] == want_picks.tolist()


def test_batch_donor_sampler_draws_the_reference_donors():
    """DonorSampler.draw consumes np.random exactly like consecutive run_original calls: the donors of the reference fixture."""
    from types import SimpleNamespace
    from hiast_b200.preprocessor import CopyPaste, DonorSampler
    spec = gi.COPY_PASTE_SPEC
    ds = gi.CopyPasteDataset(spec)
    cfg = SimpleNamespace(dataset=SimpleNamespace(source=SimpleNamespace(type='GTAV'), num_classes=spec['C']),
                          preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(selected_num_classes=spec['selected'], mode='original')))
    cp = CopyPaste(cfg, ds, gi.copy_paste_class_value(spec), device='cpu')
    np.random.seed(spec['seed'])
    got = DonorSampler(cp).draw(spec['n_run'])
    np.random.seed(spec['seed'])
    want = []
    for i in range(spec['n_run']):
        img, lbl, _ = ds.load_data(i)
        _, _, _, donors = ocp.run_original(img, lbl, cp.hard_classes, cp.class_probs, ds.get_samples_with_class(),
                                           lambda name: ds.load_data(ds.get_file_to_idx(name))[:2], spec['C'])
        assert len(donors) == 1                            # the loop stops after its first donor
        want.append(ds.get_file_to_idx(donors[0]))
    assert got == want


def test_travelling_reference_copy_runs_and_equals_the_port():
    """oracle/_ref/code (the reference's hot-path files, copied by __graft_entry__.build() where /root/reference is mounted)
    through oracle/ref.py's shim: the unmodified IASPseudoGenerator.run equals the oracle port on the same batches -- the
    CPU arm of bench.py (kind 'reference') and its fallback (kind 'port') are the same computation."""
    from oracle import ref as oref
    if oref.available() is None:
        pytest.skip('no reference copy on this box')
    spec = gi.IAS_SPECS['ias_small']
    batches = gi.ias_batches(spec)
    gen = oref.run_ias(batches, spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    o = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'], faithful=True)
    o.run(batches)
    assert np.array_equal(gen.class_threshold, o.class_threshold)
    assert np.array_equal(gen.statics_class, o.statics_class)
    assert np.array_equal(gen.class_mean_probs, o.class_mean_probs)
    assert gen.sample_stats == o.sample_stats
    assert all(np.array_equal(a, b) for a, b in zip(gen.captured, o.labels))
