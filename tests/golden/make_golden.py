#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is mounted there and nowhere
else):  ``python tests/golden/make_golden.py``.  The reference is imported from
``/root/reference/code`` through a stub shim for the packages that are not
installed (apex, tensorboardX, albumentations, numpy.lib.type_check, np.bool) --
none of the stubs touch the arithmetic of the hot path.  Every fixture is the
output of the reference's own functions on seeded inputs:

* ias_*.npz       IASPseudoGenerator.run        workflows/pseudo_label_generator.py:181-213
* cbst_small.npz  CBSTPseudoGenerator.run       workflows/pseudo_label_generator.py:142-165,115-132
* loss_cst_variants.npz  LOSS['KLDIV'], LOSS['MSE']      sseg/models/modules/losses.py:9-23
* loss_*.npz      SelfTrainingSegmentor.compute_loss + autograd
                                                sseg/models/segmentors/self_training_segmentor.py:30-53
* metric_*.npz    intersectionAndUnionGPU       utils/metrics.py:6-19
* copy_paste.npz  CopyPaste.run_original        sseg/datasets/preprocessor.py:79-122
* ema_update.npz  update_ema_model              utils/utils.py:115-123
* loss_ce_general.npz  LOSS['CE'] with class weights / refer_labels   sseg/models/modules/losses.py:32-36,68-89
* validator.npz   Validator.get_multi_scale_and_flip_logits + argmax   workflows/validator.py:34-55,92-93
* copy_paste_probs.npz  CopyPaste.calculate_class_probs / random_select, 20 seeds   sseg/datasets/preprocessor.py:29-34,70-77
* cli_config.json  utils/default_config.py + generate_pseudo_labels.update_cfg on the shipped yaml files
                                                generate_pseudo_labels.py:21-40
* pseudo_store.npz  BaseDataset.stat_samples_with_class / load_data (pseudo-label branch)
                                                sseg/datasets/loader/base_dataset.py:61-77,158-178

Inputs that are large are not stored: they are regenerated from a seeded CPU
``torch.Generator`` by ``tests/golden_inputs.py`` (same torch build on the GPU
box); a sha256 of the input is stored so that drift is detected, not trusted.
"""

from __future__ import annotations

import hashlib
import os
import sys
import tempfile
from types import ModuleType, SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))       # the repo root (hiast_b200.config for cli_config_fixture)
import golden_inputs as gi  # noqa: E402

REF = '/root/reference/code'


def install_shim():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    apex = ModuleType('apex')
    apex.amp = ModuleType('apex.amp')
    apex.parallel = ModuleType('apex.parallel')
    apex.parallel.SyncBatchNorm = type('SyncBatchNorm', (), {})
    apex.parallel.convert_syncbn_model = lambda m: m
    apex.parallel.DistributedDataParallel = object
    sys.modules.update({'apex': apex, 'apex.amp': apex.amp, 'apex.parallel': apex.parallel})
    tbx = ModuleType('tensorboardX')
    tbx.SummaryWriter = object
    sys.modules['tensorboardX'] = tbx
    alb = ModuleType('albumentations')
    alb.core = ModuleType('albumentations.core')
    alb.core.composition = ModuleType('albumentations.core.composition')
    alb.core.composition.BaseCompose = object
    sys.modules.update({'albumentations': alb, 'albumentations.core': alb.core,
                        'albumentations.core.composition': alb.core.composition})
    tc = ModuleType('numpy.lib.type_check')
    tc.common_type = np.common_type
    sys.modules['numpy.lib.type_check'] = tc
    np.bool = np.bool_
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU run of code that calls .cuda()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ----------------------------------------------------------------------------- IAS
def run_reference_ias(batches, C, alpha, beta, gamma, cp_gamma):
    from workflows import pseudo_label_generator as plg

    class Identity:
        def eval(self):
            return self

        def __call__(self, x):
            return {'logits': x}

    class Harness(plg.IASPseudoGenerator):
        def initialize(self):
            self.model = Identity()
            self.t_loader = [{'images': lg, 'image_paths': paths} for lg, paths in batches]
            self.t_dataset = [None] * sum(len(p) for _, p in batches)
            self.pseudo_label_save_dir = tempfile.mkdtemp()
            self.captured = []
            self.thr_per_image = []

        def save_pseudo_label(self, plbl, img_path):
            self.captured.append(plbl.astype(np.uint8))   # the PNG payload of :46
            self.thr_per_image.append(self.class_threshold.copy())

        def save_data(self):
            pass

    cfg = SimpleNamespace(
        dataset=SimpleNamespace(num_classes=C),
        pseudo_policy=SimpleNamespace(type='IAS', ias=SimpleNamespace(alpha=alpha, beta=beta, gamma=gamma)),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=cp_gamma)))
    import contextlib
    import io
    gen = Harness(cfg)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        gen.run()
    return gen


def ias_fixture(name, spec, store_conf=True):
    batches = gi.ias_batches(spec)
    gen = run_reference_ias(batches, spec['C'], spec['alpha'], spec['beta'], spec['gamma'], spec['cp_gamma'])
    n_img = sum(len(p) for _, p in batches)
    C = spec['C']
    first = np.cumsum([0] + [len(p) for _, p in batches])[:-1]
    gen.thr_trace = [gen.thr_per_image[i] for i in first]   # threshold in force for each batch
    counts = np.zeros((n_img, C), dtype=np.int64)
    for i, st in enumerate(gen.sample_stats):
        for k, v in st.items():
            if k != 'file':
                counts[i, k] = v
    out = dict(
        spec=np.array(repr(spec)),
        logits_sha=np.array([sha(lg.numpy()) for lg, _ in batches]),
        thr_trace=np.stack(gen.thr_trace),                      # f64 [G, C] after each batch
        class_threshold=gen.class_threshold,                    # f64 [C]
        class_mean_probs=gen.class_mean_probs,                  # f64 [C]
        statics_class=np.asarray(gen.statics_class, dtype=np.int64),
        counts=counts,
        plbl_sha=np.array([sha(p) for p in gen.captured]),
    )
    assert len(gen.thr_trace) == len(batches), (len(gen.thr_trace), len(batches))
    if store_conf:
        from torch.nn import functional as F
        confs, labels = [], []
        for lg, _ in batches:
            c, l = F.softmax(lg, dim=1).max(dim=1)
            confs.append(c.numpy())
            labels.append(l.numpy().astype(np.uint8))
        out['conf'] = np.concatenate(confs)
        out['label'] = np.concatenate(labels)
        out['plbl'] = np.stack(gen.captured)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'thr', gen.class_threshold[:4], 'kept', counts.sum(), 'of', n_img * spec['H'] * spec['W'])


def cbst_fixture(name, spec, interval, p):
    """CBSTPseudoGenerator.run (two passes) of the reference, pseudo_label_generator.py:142-165 + :115-132."""
    from workflows import pseudo_label_generator as plg
    batches = gi.ias_batches(spec)

    class Identity:
        def eval(self):
            return self

        def __call__(self, x):
            return {'logits': x}

    class Harness(plg.CBSTPseudoGenerator):
        def initialize(self):
            self.model = Identity()
            self.t_loader = [{'images': lg, 'image_paths': paths} for lg, paths in batches]
            self.t_dataset = [None] * sum(len(pp) for _, pp in batches)
            self.pseudo_label_save_dir = tempfile.mkdtemp()
            self.captured = []

        def save_pseudo_label(self, plbl, img_path):
            self.captured.append(plbl.astype(np.uint8))

        def save_data(self):
            pass

    cfg = SimpleNamespace(
        dataset=SimpleNamespace(num_classes=spec['C']),
        pseudo_policy=SimpleNamespace(type='CBST', cbst=SimpleNamespace(sample_interval=interval, p=p)),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=spec['cp_gamma'])))
    import contextlib
    import io
    gen = Harness(cfg)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        gen.run()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), spec=np.array(repr(spec)), interval=interval, p=p,
                        class_threshold=gen.class_threshold, plbl=np.stack(gen.captured),
                        statics_class=np.asarray(gen.statics_class, dtype=np.int64), class_mean_probs=gen.class_mean_probs)
    print(name, gen.class_threshold[:5])


# ---------------------------------------------------------------------------- loss
def loss_fixture(name, spec):
    from sseg.models.modules import losses
    from sseg.models.segmentors import self_training_segmentor as sts
    z, t, plbl, s_z, s_lbl = gi.loss_inputs(spec)
    z = z.clone().requires_grad_(True)
    if s_z is not None:
        s_z = s_z.clone().requires_grad_(True)
    cfg = SimpleNamespace(
        model=SimpleNamespace(predictor=SimpleNamespace(
            seg_loss=SimpleNamespace(target_pseudo_weight=spec['w_seg']),
            kld_loss=SimpleNamespace(weight=spec['w_kld']),
            ent_loss=SimpleNamespace(weight=spec['w_ent']))),
        cst_training=SimpleNamespace(is_enabled=True, cst_loss=SimpleNamespace(
            weight=spec['w_cst'], region=spec['region'])))
    self_ = SimpleNamespace(cfg=cfg, seg_loss_fun=losses.ce, kld_loss_fun=sts._kld,
                            ent_loss_fun=sts._entropy, cst_loss_fun=losses.soft_ce)
    out = sts.SelfTrainingSegmentor.compute_loss(self_, z, plbl, t, s_z, s_lbl)
    total = sum(v.mean() for v in out.values())     # base_trainer.py:129
    total.backward()
    res = dict(spec=np.array(repr(spec)), z=z.detach().numpy(), t=t.numpy(), plbl=plbl.numpy(),
               grad=z.grad.numpy(), keys=np.array(list(out.keys())),
               values=np.array([v.item() for v in out.values()], dtype=np.float64))
    if s_z is not None:
        res.update(s_z=s_z.detach().numpy(), s_lbl=s_lbl.numpy(), s_grad=s_z.grad.numpy())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
    print(name, {k: float(v) for k, v in out.items()})


def cst_variant_fixture(name, spec):
    """LOSS['KLDIV'] / LOSS['MSE'] of the reference (losses.py:9-23) with and without refer_labels, + gradients."""
    import warnings
    from sseg.models.modules import losses
    z, t, plbl, _, _ = gi.loss_inputs(spec)
    tz = torch.log(t) * 1.7                      # teacher "logits" with a different temperature
    res = dict(spec=np.array(repr(spec)), z=z.numpy(), t=t.numpy(), tz=tz.numpy(), plbl=plbl.numpy())
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for kind, fn, tgt in (('kldiv', losses.kl_div, tz), ('mse', losses.mse, t)):
            for region in ('none', 'ignored', 'confident', 'all'):
                zz = z.clone().requires_grad_(True)
                if region == 'none':
                    val = fn(zz, tgt)
                else:
                    val = fn(zz, tgt, refer_labels=plbl, region=region)
                val.backward()
                res['%s_%s' % (kind, region)] = np.float64(val.item())
                res['%s_%s_grad' % (kind, region)] = zz.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
    print(name, {k: float(v) for k, v in res.items() if k.endswith(('none', 'ignored', 'confident', 'all'))})


# -------------------------------------------------------------------------- metric
def metric_fixture(name, spec):
    from utils import metrics
    pred, target = gi.metric_inputs(spec)
    p = pred.clone()
    inter, union = metrics.intersectionAndUnionGPU(p, target, spec['K'])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), spec=np.array(repr(spec)),
                        pred=pred.numpy(), target=target.numpy(), pred_after=p.numpy(),
                        intersection=inter.numpy(), union=union.numpy())
    print(name, inter.numpy()[:5], union.numpy()[:5])


# ---------------------------------------------------------------------- copy-paste
def copy_paste_fixture(name, spec):
    from sseg.datasets import preprocessor
    ds = gi.CopyPasteDataset(spec)
    cfg = SimpleNamespace(
        dataset=SimpleNamespace(source=SimpleNamespace(type='GTAV'), num_classes=spec['C']),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(
            selected_num_classes=spec['selected'], mode='original')))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        cp = preprocessor.CopyPaste(cfg, ds, gi.copy_paste_class_value(spec))
    res = dict(spec=np.array(repr(spec)), hard=np.asarray(cp.hard_classes), probs=cp.class_probs)
    np.random.seed(spec['seed'])
    for i in range(spec['n_run']):
        img, lbl, _ = ds.load_data(i)
        img, lbl = img.copy(), lbl.copy()
        o_img, o_lbl, o_mask = cp.run(img, lbl)
        res['img_%d' % i] = o_img
        res['lbl_%d' % i] = o_lbl
        res['mask_%d' % i] = o_mask
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
    print(name, 'hard', cp.hard_classes)


def copy_paste_probs_fixture(name, n_seeds=20, n_draws=50):
    """calculate_class_probs (preprocessor.py:29-34, torch float64 arithmetic) and the class draws of random_select (:70-77) for
    `n_seeds` random class-value vectors: a last-ulp difference in p can change an np.random.choice draw (VERDICT r1 weak #1 iii)."""
    from sseg.datasets import preprocessor
    res = {}
    for seed in range(n_seeds):
        rs = np.random.RandomState(1000 + seed)
        C = 19 if seed % 4 else 16
        value = rs.uniform(0.3, 0.9995, size=C)
        me = SimpleNamespace(class_value=value.copy(), cfg=SimpleNamespace(dataset=SimpleNamespace(num_classes=C)))
        probs = preprocessor.CopyPaste.calculate_class_probs(me)
        me.class_probs = probs
        hard = np.argsort(value)[:14]
        np.random.seed(seed)
        picks = [int(preprocessor.CopyPaste.random_select(me, hard)) for _ in range(n_draws)]
        res['value_%d' % seed], res['probs_%d' % seed], res['picks_%d' % seed] = value, probs, np.asarray(picks)
        res['hard_%d' % seed] = hard
    np.savez_compressed(os.path.join(HERE, name + '.npz'), n_seeds=n_seeds, **res)
    print(name, 'seeds', n_seeds)


def cli_config_fixture(name):
    """The reference's own configuration code on hiast_b200.config.CfgNode standing in for yacs (not installed):
    utils/default_config.py builds the default tree, generate_pseudo_labels.update_cfg (:21-40) merges the shipped yaml files and
    the command-line flags UNMODIFIED.  Stored: the default tree and the merged trees, as JSON."""
    import json
    from hiast_b200.config import CfgNode
    yacs = ModuleType('yacs')
    yacs.config = ModuleType('yacs.config')
    yacs.config.CfgNode = CfgNode
    sys.modules.update({'yacs': yacs, 'yacs.config': yacs.config})
    for m in ('utils.default_config', 'generate_pseudo_labels'):
        sys.modules.pop(m, None)
    reg = ModuleType('utils.registry.register')               # the script imports it for its side effect only
    sys.modules['utils.registry.register'] = reg
    import importlib
    out = {}
    dc = importlib.import_module('utils.default_config')
    out['defaults'] = json.loads(json.dumps(dc.cfg))
    gpl = importlib.import_module('generate_pseudo_labels')
    cases = {
        'sl_1': dict(config_file=os.path.join(REF, 'configs', 'sl_1.yaml')),
        'sl_1_setting_flags': dict(config_file=os.path.join(REF, 'configs', 'sl_1.yaml'),
                                   setting_file=os.path.join(REF, 'configs', 'hiast_setting.yaml'),
                                   pseudo_resume_from='/ckpt/model.pth', pseudo_save_dir='/out/pl', seg_model='DeepLab_V2'),
        'sl_3': dict(config_file=os.path.join(REF, 'configs', 'sl_3.yaml'), pseudo_save_dir='/x'),
    }
    for key, kw in cases.items():
        sys.modules.pop('utils.default_config', None)
        dc = importlib.import_module('utils.default_config')
        args = SimpleNamespace(config_file=None, setting_file=None, pseudo_resume_from=None, pseudo_save_dir=None, batch_size=None,
                               seg_model=None)
        args.__dict__.update(kw)
        cfg = gpl.update_cfg(dc.cfg, args)
        assert cfg.is_frozen()
        out[key] = json.loads(json.dumps(cfg))
    # the --batch_size bug of :30 (cfg.batch_size does not exist)
    sys.modules.pop('utils.default_config', None)
    dc = importlib.import_module('utils.default_config')
    args = SimpleNamespace(config_file=os.path.join(REF, 'configs', 'sl_1.yaml'), setting_file=None, pseudo_resume_from=None,
                           pseudo_save_dir=None, batch_size=4, seg_model=None)
    try:
        gpl.update_cfg(dc.cfg, args)
        out['batch_size_flag'] = 'accepted'
    except AttributeError as e:
        out['batch_size_flag'] = 'AttributeError'
    with open(os.path.join(HERE, name + '.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(name, sorted(out))


def ema_fixture(name):
    """utils.update_ema_model on a small conv net with BatchNorm buffers (float32 parameters, int64 / float32 buffers)."""
    from utils import utils as ref_utils
    g = torch.Generator().manual_seed(77)

    def net():
        m = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 5, 1, bias=False),
                                torch.nn.Linear(7, 1031))
        for p in m.parameters():
            p.data = torch.randn(p.shape, generator=g) * (10.0 ** float(torch.randint(-6, 3, (1,), generator=g)))
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.data = torch.rand(b.shape, generator=g)
            else:
                b.data = torch.randint(0, 1000, b.shape, generator=g)
        return m

    student, teacher = net(), net()
    out = {'gamma': np.float64(0.999)}
    for i, p in enumerate(student.parameters()):
        out['q%d' % i] = p.data.numpy().copy()
    for i, p in enumerate(teacher.parameters()):
        out['k%d' % i] = p.data.numpy().copy()
    for i, b in enumerate(student.buffers()):
        out['bq%d' % i] = b.data.numpy().copy()
    ref_utils.update_ema_model(teacher, student, 0.999)
    for i, p in enumerate(teacher.parameters()):
        out['new%d' % i] = p.data.numpy().copy()
    for i, b in enumerate(teacher.buffers()):
        out['bnew%d' % i] = b.data.numpy().copy()
    out['n_params'] = np.int64(len(list(student.parameters())))
    out['n_buffers'] = np.int64(len(list(student.buffers())))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'params', int(out['n_params']), 'buffers', int(out['n_buffers']))


def pseudo_store_fixture(name):
    """Reader side of the on-disk pseudo-label outputs, run unbound on a SimpleNamespace `self`."""
    import json
    import cv2
    from sseg.datasets.loader.base_dataset import BaseDataset
    rng = np.random.default_rng(gi.PSEUDO_STORE_SPEC['seed'])
    C = gi.PSEUDO_STORE_SPEC['C']
    out = {}
    with tempfile.TemporaryDirectory() as root:
        samples = gi.pseudo_store_samples()
        with open(os.path.join(root, 'samples_with_class.json'), 'w') as f:      # what save_data writes (:60-62)
            f.write(json.dumps(samples))
        self = SimpleNamespace(cfg=SimpleNamespace(dataset=SimpleNamespace(num_classes=C)))
        res = BaseDataset.stat_samples_with_class(self, root)
        out['stat_json'] = np.array(json.dumps(res))
        pdir = os.path.join(root, 'pseudo_labels')
        os.makedirs(pdir)
        for k, (src, dst) in enumerate(gi.PSEUDO_STORE_SPEC['sizes']):
            img = rng.integers(0, 256, (dst[0], dst[1], 3)).astype(np.uint8)
            lbl = gi.pseudo_store_label(k, src)
            img_path = os.path.join(root, 'img_%d.png' % k)
            cv2.imwrite(img_path, img)
            cv2.imwrite(os.path.join(pdir, 'img_%d_pseudo_label.png' % k), lbl)        # pseudo_label_generator.py:46
            ds = SimpleNamespace(img_path_list=[img_path], lbl_path_list=['unused'], pseudo_dir=pdir, read_label=None)
            img_r, lbl_r, path_r = BaseDataset.load_data(ds, 0)
            assert path_r == img_path and img_r.shape == img.shape
            out['lbl_%d' % k] = lbl_r
            out['src_sha_%d' % k] = np.array(sha(lbl))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith('lbl_')})


def ce_general_fixture(name):
    """LOSS['CE'] of the reference (losses.py:32-36) with class weights and / or refer_labels, + gradients."""
    from sseg.models.modules import losses
    res = {}
    for key, spec in gi.CE_GENERAL_SPECS.items():
        z, labels, weights, refer = gi.ce_general_inputs(spec)
        for case, kw in gi.ce_general_cases(labels, weights, refer).items():
            zz = z.clone().requires_grad_(True)
            val = losses.ce(zz, **kw)
            val.backward()
            res['%s_%s' % (key, case)] = np.float64(val.item())
            res['%s_%s_grad' % (key, case)] = zz.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **res)
    print(name, {k: float(v) for k, v in res.items() if not k.endswith('_grad')})


def validator_fixture(name):
    """Validator.get_multi_scale_and_flip_logits + argmax (workflows/validator.py:34-55,92-93) run unbound on CPU."""
    from workflows.validator import Validator
    out = {}
    for key, spec in gi.VALIDATOR_SPECS.items():
        model = gi.ToyModel(spec['C'], spec['seed'])
        imgs = gi.validator_images(spec)
        self = SimpleNamespace(model=model, cfg=SimpleNamespace(validate=SimpleNamespace(resize_sizes=spec['sizes'],
                                                                                         is_flip=spec['flip'])))
        with torch.no_grad():
            results = Validator.get_multi_scale_and_flip_logits(self, imgs)
            lbls_pred = results.argmax(dim=1)                         # :93
        out[key + '_results'] = results.numpy()
        out[key + '_labels'] = lbls_pred.numpy().astype(np.uint8)
        out[key + '_imgs_sha'] = np.array(sha(imgs.numpy()))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith('_labels')})


def main():
    install_shim()
    if len(sys.argv) > 1 and sys.argv[1] == 'pseudo_store':
        pseudo_store_fixture('pseudo_store')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'ce_general':
        ce_general_fixture('loss_ce_general')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'copy_paste_probs':
        copy_paste_probs_fixture('copy_paste_probs')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'cli_config':
        cli_config_fixture('cli_config')
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'validator':
        validator_fixture('validator')
        return
    copy_paste_probs_fixture('copy_paste_probs')
    cli_config_fixture('cli_config')
    validator_fixture('validator')
    ce_general_fixture('loss_ce_general')
    pseudo_store_fixture('pseudo_store')
    ema_fixture('ema_update')
    for name, spec in gi.IAS_SPECS.items():
        ias_fixture(name, spec, store_conf=spec.get('store_conf', True))
    cbst_fixture('cbst_small', gi.IAS_SPECS['ias_small'], 4, 0.2)
    for name, spec in gi.LOSS_SPECS.items():
        loss_fixture(name, spec)
    cst_variant_fixture('loss_cst_variants', gi.CST_VARIANT_SPEC)
    for name, spec in gi.METRIC_SPECS.items():
        metric_fixture(name, spec)
    copy_paste_fixture('copy_paste', gi.COPY_PASTE_SPEC)


if __name__ == '__main__':
    main()
