"""The reference-facing generator API (PSEUDO_POLICY[...](cfg).run()) and the sharded driver on the GPU."""

import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import ias as oias

pytestmark = pytest.mark.gpu


class Identity:
    def eval(self):
        return self

    def __call__(self, x):
        return {'logits': x}


def make_cfg(spec, ptype='IAS'):
    return SimpleNamespace(
        dataset=SimpleNamespace(num_classes=spec['C']),
        pseudo_policy=SimpleNamespace(type=ptype, batch_size=spec['B'], ct=SimpleNamespace(threshold=0.6),
                                      ias=SimpleNamespace(alpha=spec['alpha'], beta=spec['beta'], gamma=spec['gamma'])),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=spec['cp_gamma'])))


def loader_of(batches):
    return [{'images': lg, 'image_paths': p} for lg, p in batches]


@pytest.mark.parametrize('name', ['ias_small', 'ias_c7'])
@pytest.mark.parametrize('window_batches,png_mode', [(1, 'device'), (3, 'device'), (8, 'device'), (3, 'host')])
def test_ias_generator_run_matches_oracle_and_writes_reference_files(name, window_batches, png_mode, tmp_path):
    import cv2
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = gi.IAS_SPECS[name]
    batches = gi.ias_batches(spec)
    save_dir = str(tmp_path / 'run' / 'pseudo_labels')
    gen = PSEUDO_POLICY['IAS'](make_cfg(spec), model=Identity(), loader=loader_of(batches), dataset_len=spec['N'],
                               save_dir=save_dir, window_batches=window_batches, png=png_mode)
    gen.run()
    oracle = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    assert np.array_equal(gen.class_threshold, oracle.class_threshold)
    assert np.array_equal(np.concatenate(gen.threshold_trace), np.stack(oracle.threshold_trace))
    assert np.array_equal(gen.statics_class, oracle.statics_class)
    np.testing.assert_allclose(gen.class_mean_probs, oracle.class_mean_probs, rtol=1e-6)
    assert gen.sample_stats == oracle.sample_stats
    assert gen.samples_class == oracle.samples_class
    assert gen.pow_rounding_certified
    paths = [p for _, ps in batches for p in ps]
    for i, p in enumerate(paths):                                     # PNG payload == oracle label map
        file = os.path.join(save_dir, os.path.splitext(p)[0] + '_pseudo_label.png')
        png = cv2.imread(file, cv2.IMREAD_UNCHANGED)
        assert np.array_equal(png, oracle.labels[i])
        if png_mode == 'device':                                      # files written by hiast_png_encode: exact bytes
            from oracle import png as opng
            assert open(file, 'rb').read() == opng.encode_png(oracle.labels[i].astype(np.uint8))
    root = os.path.join(save_dir, '..')                               # save_data: same files as the reference
    assert np.array_equal(np.load(os.path.join(root, 'class_threshold.npy')), oracle.class_threshold)
    assert np.array_equal(np.load(os.path.join(root, 'statics_class.npy')), oracle.statics_class)
    assert json.load(open(os.path.join(root, 'sample_class_stats.json'))) == \
        json.loads(json.dumps(oracle.sample_stats))
    assert json.load(open(os.path.join(root, 'samples_with_class.json'))) == \
        json.loads(json.dumps(oracle.samples_class))
    # second run() is a no-op once the directory is full (:182-183)
    gen2 = PSEUDO_POLICY['IAS'].__new__(PSEUDO_POLICY['IAS'])
    assert len(os.listdir(save_dir)) == spec['N']


@pytest.mark.parametrize('ptype', ['CT', 'NT'])
def test_constant_and_no_threshold_policies(ptype, tmp_path):
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = gi.IAS_SPECS['ias_small']
    batches = gi.ias_batches(spec)
    captured = {}

    class Gen(PSEUDO_POLICY[ptype]):
        def save_pseudo_label(self, plbl, img_path):
            captured[img_path] = plbl.copy()

        def save_data(self):
            pass

    gen = Gen(make_cfg(spec, ptype), model=Identity(), loader=loader_of(batches), dataset_len=spec['N'],
              save_dir=str(tmp_path / 'p'), window_batches=2)
    gen.run()
    thr = None if ptype == 'NT' else 0.6 * np.ones(spec['C'])
    stats = np.zeros(spec['C'], dtype=np.int64)
    for lg, paths in batches:
        conf, label = oias.softmax_max(lg.cuda())
        for k, p in enumerate(paths):
            want = label[k] if thr is None else oias.select_confident(conf[k], label[k], thr)
            assert np.array_equal(captured[p], want.astype(np.uint8))
            stats += np.bincount(want[want != 255], minlength=spec['C'])
    assert np.array_equal(gen.statics_class, stats)


def test_reference_method_surface(tmp_path):
    """get_ias_threshold / select_and_save_confident_label keep the reference's signatures and results."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = gi.IAS_SPECS['ias_small']
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ias_small.npz'))
    C = spec['C']
    gen = PSEUDO_POLICY['IAS'](make_cfg(spec), model=Identity(), loader=[], dataset_len=0,
                               save_dir=str(tmp_path / 'pl'), png_workers=0)
    conf, label = gold['conf'][:2], gold['label'][:2].astype(np.int64)
    thr = 0.9 * np.ones(C)
    d = {c: [thr[c]] + list(conf[label == c].astype(np.float16)) for c in range(C)}     # the reference's dict (:198-201)
    temp = gen.get_ias_threshold(d, C, spec['alpha'], thr, spec['gamma'])
    want = oias.ias_quantile_thresholds(conf, label, thr, C, spec['alpha'], spec['gamma'])
    assert temp.dtype == np.float32 and np.array_equal(temp, want)
    gen.class_threshold = oias.ias_ema_update(thr, temp, spec['beta'])
    assert np.array_equal(gen.class_threshold, gold['thr_trace'][0])
    last = gen.select_and_save_confident_label(conf, label, ['a.png', 'b.png'])
    assert np.array_equal(last.astype(np.uint8), gold['plbl'][1])
    assert np.array_equal(gen.statics_class, gold['counts'][:2].sum(0))
    assert sorted(os.listdir(str(tmp_path / 'pl'))) == ['a_pseudo_label.png', 'b_pseudo_label.png']


@pytest.mark.parametrize('slots,reserve', [(2, 0), (3, 0), (3, 8), (3, 40)])
def test_windowed_ring_driver_single_rank_equals_oracle(slots, reserve):
    """ShardedIAS with world_size 1 (the bench path), windows of 2 groups: two slots (outputs behind the chain), three slots
    (chain of window j-1 beside the outputs of j-2) and the CONCURRENT schedule (reserve > 0: phase A leaves SMs free, scan
    and phase C of window j run there on the chain stream while phase A of window j+1 runs)."""
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.sharded import ShardedIAS, window_images
    spec = gi.IAS_SPECS['ias_small']
    batches = gi.ias_batches(spec)
    logits = torch.cat([lg for lg, _ in batches]).cuda()
    window = 2 * spec['B']
    eng = IASEngine(spec['C'], spec['H'], spec['W'], spec['B'], spec['alpha'], spec['beta'], spec['gamma'],
                    spec['cp_gamma'], slots * window)
    eng.reserve_sms = reserve
    got = {}

    def on_window(w, plbl, counts, thr_groups):
        got[w] = (plbl.cpu().numpy(), thr_groups.cpu().numpy())

    def window_logits(w):
        i0, n = window_images(w, window, spec['N'])
        return logits[i0:i0 + n]

    thr, mean, statics = ShardedIAS(eng, window, spec['N'], 0, 1).run(window_logits, on_window)
    torch.cuda.synchronize()
    oracle = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    assert np.array_equal(thr.cpu().numpy(), oracle.class_threshold)
    assert np.array_equal(statics.cpu().numpy(), oracle.statics_class)
    np.testing.assert_allclose(mean.cpu().numpy(), oracle.class_mean_probs, rtol=1e-6)
    assert np.array_equal(np.concatenate([got[w][0] for w in sorted(got)]), np.stack(oracle.labels))
    assert np.array_equal(np.concatenate([got[w][1] for w in sorted(got)]), np.stack(oracle.threshold_trace))


def test_config4_synthia_16_class_ias_plus_copy_paste():
    """BASELINE.json configs[3] on one GPU: 19-channel tensors whose classes {9,14,16} are never predicted, IAS through
    the windowed driver, then hard-aware copy-paste on uint8 images with labels = IAS output and donor (i*7919+13) mod N."""
    from hiast_b200.ias_engine import IASEngine
    from hiast_b200.preprocessor import CopyPaste
    from hiast_b200.sharded import ShardedIAS, window_images
    from oracle import copy_paste as ocp
    spec = dict(C=19, H=32, W=64, N=10, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99, seed=77, dist='peaked',
                absent=(9, 14, 16))
    batches = gi.ias_batches(spec)
    logits = torch.cat([lg for lg, _ in batches]).cuda()
    window = 4
    eng = IASEngine(19, 32, 64, 2, 0.5, 0.9, 8.0, 0.99, 2 * window)
    labels = {}
    def window_logits(w):
        i0, n_w = window_images(w, window, spec['N'])
        return logits[i0:i0 + n_w]

    def on_window(w, plbl_w, counts, thr_groups):
        labels[w] = plbl_w.clone()

    thr, mean, statics = ShardedIAS(eng, window, spec['N'], 0, 1).run(window_logits, on_window)
    oracle = oias.IASOracle(19, 0.5, 0.9, 8.0, 0.99)
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    plbl = torch.cat([labels[w] for w in sorted(labels)])
    assert np.array_equal(plbl.cpu().numpy(), np.stack(oracle.labels))
    assert np.array_equal(thr.cpu().numpy(), oracle.class_threshold)
    assert statics[[9, 14, 16]].sum().item() == 0 and (mean[[9, 14, 16]] == 0).all()
    # copy-paste with the SYNTHIA class set: ignored classes get p = 0 and never enter the hard set
    cfg = SimpleNamespace(dataset=SimpleNamespace(source=SimpleNamespace(type='SYNTHIA'), num_classes=19),
                          preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(selected_num_classes=10, mode='original')))

    class DS:
        def get_samples_with_class(self):
            return {c: ['x'] for c in range(19)}

    class_value = mean.cpu().numpy().copy()
    class_value[class_value == 0] = 1.0
    cp = CopyPaste(cfg, DS(), class_value)
    assert not set(int(c) for c in cp.hard_classes) & {9, 14, 16}
    n = spec['N']
    rs = np.random.RandomState(1)
    imgs = torch.from_numpy(rs.randint(0, 256, size=(n, 32, 64, 3)).astype(np.uint8)).cuda()
    donor_imgs, donor_lbls = imgs.clone(), plbl.clone()
    donor_index = torch.tensor([(i * 7919 + 13) % n for i in range(n)], dtype=torch.int32, device='cuda')
    out_img, out_lbl, out_mask = cp.run_batch(imgs.clone(), plbl.clone(), donor_imgs, donor_lbls, donor_index)
    for i in range(n):
        d = int(donor_index[i])
        w_img, w_lbl = imgs[i].cpu().numpy().copy(), plbl[i].cpu().numpy().copy()
        w_mask = np.full_like(w_lbl, 255)
        ocp.paste(w_img, w_lbl, w_mask, donor_imgs[d].cpu().numpy(), donor_lbls[d].cpu().numpy(), cp.hard_classes)
        assert np.array_equal(out_img[i].cpu().numpy(), w_img)
        assert np.array_equal(out_lbl[i].cpu().numpy(), w_lbl)
        assert np.array_equal(out_mask[i].cpu().numpy(), w_mask)


def test_config5_full_round_example_small():
    """BASELINE.json configs[4] at toy size: DeepLabv2-ResNet101 forward -> IAS -> confusion matrix, via the example."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'examples', 'full_round.py'), '--images', '5', '--height', '65',
                          '--width', '129', '--window', '2'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res['images'] == 5 and 0.0 <= res['miou'] <= 1.0 and res['pow_rounding_certified']


def test_cbst_policy_matches_oracle_and_reference_fixture(tmp_path):
    """PSEUDO_POLICY['CBST'] (pseudo_label_generator.py:142-165): every-k-th raster-order sampling, quantile thresholds,
    second pass with the constant thresholds.  Bit-exact vs the oracle (float64 quantile = the reference's pinned
    numpy 1.19 semantic, see oracle.ias.cbst_thresholds) on CUDA softmax; within one fp16 step of the fixture, which
    numpy 2.3.5 computed with a float16 virtual index."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = gi.IAS_SPECS['ias_small']
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'cbst_small.npz'))
    batches = gi.ias_batches(spec)
    captured = []

    class Gen(PSEUDO_POLICY['CBST']):
        def save_pseudo_label(self, plbl, img_path):
            captured.append(plbl.copy())

        def save_data(self):
            pass

    cfg = make_cfg(spec, 'CBST')
    cfg.pseudo_policy.cbst = SimpleNamespace(sample_interval=int(gold['interval']), p=float(gold['p']))
    gen = Gen(cfg, model=Identity(), loader=loader_of(batches), dataset_len=spec['N'], save_dir=str(tmp_path / 'p'),
              window_batches=2)
    gen.run()
    cl = [oias.softmax_max(lg.cuda()) for lg, _ in batches]
    thr = oias.cbst_thresholds(cl, spec['C'], int(gold['interval']), float(gold['p']))
    assert np.array_equal(gen.class_threshold, thr)
    want = [oias.select_confident(c[k], l[k], thr).astype(np.uint8) for c, l in cl for k in range(len(c))]
    assert np.array_equal(np.stack(captured), np.stack(want))
    assert np.abs(gen.class_threshold - gold['class_threshold']).max() < 2e-3


def test_cbst_sampling_full_resolution_properties():
    """19x1024x2048, batch 2: the sampled histogram holds ceil(n_c / k) samples per class and batch."""
    from hiast_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(5)
    logits = torch.randn(4, 19, 1024, 2048, generator=g, device='cuda') * 3
    conf, label, _ = ops.ias_softmax_hist(logits, 2)
    key_lo = ops.ias_key_lo(19)
    hist = torch.zeros((19, ops.ias_row_stride(key_lo)), dtype=torch.int32, device='cuda')
    ops.cbst_sample_hist(conf, label, 19, 2, 4, key_lo, hist)
    want = torch.zeros(19, dtype=torch.int64, device='cuda')
    for b in range(2):
        n_c = torch.bincount(label[2 * b:2 * b + 2].flatten().long(), minlength=19)
        want += (n_c + 3) // 4
    assert torch.equal(hist.sum(1).long(), want)
    # the samples of class c are exactly conf[label == c][::4] of each batch (raster order)
    c = 7
    vals = torch.cat([conf[2 * b:2 * b + 2][label[2 * b:2 * b + 2] == c][::4] for b in range(2)])
    keys = vals.half().view(torch.int16).long() - key_lo
    assert torch.equal(hist[c].long(), torch.bincount(keys, minlength=hist.shape[1]))


@pytest.mark.parametrize('ptype', ['IAS', 'CT', 'CBST'])
def test_generators_from_stride8_logits_equal_the_interpolated_path(ptype, tmp_path):
    """SURVEY 8f rank 1 through the reference-facing API: a model that returns its stride-8 logits ('logits_lr') gives the
    same thresholds / statistics / PNGs as the reference's F.interpolate(..., align_corners=True) inside the model."""
    import cv2
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = dict(C=19, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99)
    H, W, h, w, N = 64, 128, 9, 17, 6
    g = torch.Generator().manual_seed(5)
    imgs = [torch.randn(2, 19, h, w, generator=g) * 4 for _ in range(N // 2)]      # stand-ins: the "image" IS the low-res logit map

    class Full:
        def eval(self):
            return self

        def __call__(self, x):
            return {'logits': torch.nn.functional.interpolate(x, size=(H, W), mode='bilinear', align_corners=True)}

    class Low(Full):
        def __call__(self, x):
            return {'logits_lr': x, 'size': (H, W)}

    res = {}
    for name, model in (('full', Full()), ('low', Low())):
        cfg = make_cfg(spec, ptype)
        cfg.pseudo_policy.cbst = SimpleNamespace(sample_interval=4, p=0.3) if ptype == 'CBST' else None
        loader = [{'images': x, 'image_paths': ['%s_%d.png' % (name, 2 * i + k) for k in range(2)]} for i, x in enumerate(imgs)]
        save_dir = str(tmp_path / name / 'pseudo_labels')
        gen = PSEUDO_POLICY[ptype](cfg, model=model, loader=loader, dataset_len=N, save_dir=save_dir, window_batches=2)
        gen.run()
        pngs = [cv2.imread(os.path.join(save_dir, '%s_%d_pseudo_label.png' % (name, i)), cv2.IMREAD_UNCHANGED) for i in range(N)]
        res[name] = (np.asarray(gen.class_threshold), np.asarray(gen.statics_class), np.asarray(gen.class_mean_probs), pngs)
    assert np.array_equal(res['full'][0], res['low'][0])
    assert np.array_equal(res['full'][1], res['low'][1])
    assert np.array_equal(res['full'][2], res['low'][2])
    for a, b in zip(res['full'][3], res['low'][3]):
        assert a is not None and np.array_equal(a, b)


@pytest.mark.parametrize('prefetch', [0, 1, None])
def test_h2d_prefetch_thread_does_not_change_results(prefetch, tmp_path):
    """The producer thread that issues the host-to-device copies ahead (prefetch > 0 / None = auto) is pure plumbing:
    thresholds, statistics and files equal the inline path and the oracle; a loader error surfaces in run()."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    spec = gi.IAS_SPECS['ias_small']
    batches = gi.ias_batches(spec)
    pinned = [(lg.pin_memory(), p) for lg, p in batches]
    save_dir = str(tmp_path / 'run' / 'pseudo_labels')
    gen = PSEUDO_POLICY['IAS'](make_cfg(spec), model=Identity(), loader=loader_of(pinned), dataset_len=spec['N'],
                               save_dir=save_dir, window_batches=2, prefetch=prefetch)
    gen.run()
    oracle = oias.IASOracle(spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    oracle.run([(lg.cuda(), p) for lg, p in batches])
    assert np.array_equal(gen.class_threshold, oracle.class_threshold)
    assert np.array_equal(gen.statics_class, oracle.statics_class)
    assert gen.sample_stats == oracle.sample_stats
    assert len(os.listdir(save_dir)) == spec['N']

    def broken():
        yield {'images': pinned[0][0], 'image_paths': pinned[0][1]}
        raise OSError('disk gone')

    gen2 = PSEUDO_POLICY['IAS'](make_cfg(spec), model=Identity(), loader=broken(), dataset_len=spec['N'],
                                save_dir=str(tmp_path / 'run2' / 'pseudo_labels'), window_batches=2, prefetch=prefetch)
    with pytest.raises(OSError):
        gen2.run()


def test_cli_script_end_to_end(tmp_path, monkeypatch):
    """the reference's script surface (generate_pseudo_labels.py:8-48) on a temp yaml -- six flags, yaml merge,
    ONE-argument PSEUDO_POLICY[type](cfg) resolving the model through MODEL and the target set through DATASET -- and the
    files of pseudo_label_generator.py:43-62 on disk with the oracle's contents."""
    import cv2
    import hiast_b200
    from hiast_b200 import cli
    from hiast_b200.registry import DATASET, MODEL
    hiast_b200.register_all()
    C, H, W, N = 7, 24, 40, 9

    class ToySegmentor(torch.nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.head = torch.nn.Conv2d(3, cfg.dataset.num_classes, 1)

        def forward(self, x):
            return {'logits': self.head(x) * 6}

    class ToyTarget(torch.utils.data.Dataset):
        def __init__(self, cfg, json_path, image_dir, aug_type=None, num_classes=None):
            g = torch.Generator().manual_seed(3)
            self.imgs = torch.randn(N, 3, H, W, generator=g)

        def __len__(self):
            return N

        def __getitem__(self, i):
            return {'images': self.imgs[i], 'image_paths': 'city/img_%02d.png' % i}

    monkeypatch.setitem(MODEL, 'ToySegmentor', ToySegmentor)
    monkeypatch.setitem(DATASET, 'ToyTarget', ToyTarget)
    torch.manual_seed(11)
    ref_model = ToySegmentor(SimpleNamespace(dataset=SimpleNamespace(num_classes=C)))
    ckpt = str(tmp_path / 'ckpt.pth')
    torch.save(ref_model.state_dict(), ckpt)
    cfg_file = tmp_path / 'sl.yaml'
    cfg_file.write_text("model:\n  type: 'ToySegmentor'\ndataset:\n  num_classes: %d\n  num_workers: 0\n  target:\n    type: 'ToyTarget'\n"
                        "    json_path: 't.json'\n    image_dir: 't'\npseudo_policy:\n  batch_size: 2\n  resize_size: [ %d, %d ]\n"
                        "  type: 'IAS'\n  ias:\n    alpha: 0.5\n    beta: 0.9\n    gamma: 8.0\n" % (C, H, W))
    save_dir = str(tmp_path / 'run' / 'pseudo_labels')
    torch.manual_seed(5)                                    # the reference's DataLoader shuffles (:36): pin the order
    gen = cli.main(['--config_file', str(cfg_file), '--pseudo_resume_from', ckpt, '--pseudo_save_dir', save_dir])
    files = sorted(os.listdir(save_dir))
    assert files == ['img_%02d_pseudo_label.png' % i for i in range(N)]
    # the same pass through the oracle, in the order the generator saw the images
    order = [row['file'] for row in gen.sample_stats]
    assert sorted(order) == ['city/img_%02d.png' % i for i in range(N)]
    ds = ToyTarget(None, None, None)
    idx = [int(p[-6:-4]) for p in order]
    ref_model = ref_model.cuda()
    with torch.no_grad():                                   # batch by batch like the generator (cuDNN picks per shape)
        batches = [(ref_model(ds.imgs[idx[i:i + 2]].cuda())['logits'], order[i:i + 2]) for i in range(0, N, 2)]
    oracle = oias.IASOracle(C, 0.5, 0.9, 8.0, 0.99)
    oracle.run(batches)
    assert np.array_equal(gen.class_threshold, oracle.class_threshold)
    assert gen.sample_stats == oracle.sample_stats
    for i, p in enumerate(order):
        png = cv2.imread(os.path.join(save_dir, os.path.splitext(os.path.basename(p))[0] + '_pseudo_label.png'), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(png, oracle.labels[i])
    root = os.path.join(save_dir, '..')
    assert np.array_equal(np.load(os.path.join(root, 'class_threshold.npy')), oracle.class_threshold)
    assert np.array_equal(np.load(os.path.join(root, 'statics_class.npy')), oracle.statics_class)
    np.testing.assert_allclose(np.load(os.path.join(root, 'class_mean_probabilities.npy')), oracle.class_mean_probs, rtol=1e-6)
    assert json.load(open(os.path.join(root, 'samples_with_class.json'))) == json.loads(json.dumps(oracle.samples_class))
    assert json.load(open(os.path.join(root, 'sample_class_stats.json'))) == json.loads(json.dumps(oracle.sample_stats))


def test_unsupported_lowres_shape_falls_back_to_interpolate(tmp_path):
    """the fused up-sampling kernel covers C in {16, 19} and W % 4 == 0; any other stride-8 input (here the
    reference's 9-class Cityscapes->Oxford setting, odd width) takes F.interpolate + the generic phase A, same results."""
    import hiast_b200
    hiast_b200.register_all()
    from hiast_b200 import PSEUDO_POLICY
    C, H, W, N, B = 9, 40, 66, 5, 2
    g = torch.Generator().manual_seed(8)
    lrs = [torch.randn(min(B, N - i), C, 6, 9, generator=g) * 4 for i in range(0, N, B)]
    paths = [['im%d.png' % (i + k) for k in range(len(lrs[i // B]))] for i in range(0, N, B)]

    class LowRes:
        def __call__(self, x):
            return {'logits_lr': x, 'size': (H, W)}

    spec = dict(C=C, B=B, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99)
    gen = PSEUDO_POLICY['IAS'](make_cfg(spec), model=LowRes(), loader=[{'images': a, 'image_paths': p} for a, p in zip(lrs, paths)],
                               dataset_len=N, save_dir=str(tmp_path / 'r' / 'pl'), window_batches=2)
    gen.run()
    full = [torch.nn.functional.interpolate(a.cuda(), size=(H, W), mode='bilinear', align_corners=True) for a in lrs]
    oracle = oias.IASOracle(C, 0.5, 0.9, 8.0, 0.99)
    oracle.run(list(zip(full, paths)))
    assert np.array_equal(gen.class_threshold, oracle.class_threshold)
    assert gen.sample_stats == oracle.sample_stats
