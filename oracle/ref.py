"""The UNMODIFIED reference as the CPU arm: build recipe + import shim + IAS harness.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

``build()`` -- called by ``__graft_entry__.build()`` -- copies the reference's Python sources from ``/root/reference/code``
(present in the build container only) into ``oracle/_ref/code``.  ``oracle/_ref/`` is git-ignored (no reference source ever
enters the history) but travels with the repo snapshot to the GPU box, the way ``baseline/_ref`` does for pip-installable
references.  The reference is pure Python (no build step): the copy IS the build.

``install_shim()`` makes ``workflows.pseudo_label_generator`` importable: stub modules for the packages that are not installed
(apex, tensorboardX, albumentations, ``numpy.lib.type_check``, ``np.bool``) -- none of them touches the arithmetic of the hot
path -- and ``Tensor.cuda()`` as the identity for the CPU run.  ``run_ias`` executes ``IASPseudoGenerator.run``
(workflows/pseudo_label_generator.py:181-213) on caller-provided logit batches: the model is the identity on the logits and
``save_pseudo_label`` captures the label map instead of calling ``cv2.imwrite`` (the PNG encode / write is excluded, stated
wherever the number is quoted).
"""

from __future__ import annotations

import contextlib
import io
import os
import shutil
import sys
import tempfile
from types import ModuleType, SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = '/root/reference/code'
REF_DIR = os.path.join(HERE, '_ref', 'code')


# The hot-path files of SURVEY.md section 8(a) + the registry they register into + the script / config of the call surface.
# Everything else of the reference (trainers, datasets, backbone, utils/utils.py) is stubbed by install_shim.
FILES = [
    'workflows/__init__.py', 'workflows/pseudo_label_generator.py',
    'sseg/__init__.py', 'sseg/models/__init__.py', 'sseg/models/segmentors/__init__.py',
    'sseg/models/segmentors/self_training_segmentor.py',
    'sseg/models/modules/__init__.py', 'sseg/models/modules/losses.py',
    'sseg/datasets/__init__.py', 'sseg/datasets/preprocessor.py',
    'utils/__init__.py', 'utils/metrics.py', 'utils/default_config.py',
    'utils/registry/__init__.py', 'utils/registry/registry.py', 'utils/registry/registries.py',
    'generate_pseudo_labels.py',
]


def build(verbose=False):
    """Copy the hot-path files of the reference into oracle/_ref/code.  Returns the directory, or None when the reference is
    not mounted and no earlier copy exists (GPU box: the copy made in the build container is used)."""
    if not os.path.isdir(SOURCE):
        return REF_DIR if os.path.isdir(REF_DIR) else None
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, rel), dst)
    if verbose:
        print('oracle/_ref: %d reference files copied from %s' % (len(FILES), SOURCE))
    return REF_DIR


def available():
    """Directory of the reference sources to import from (the travelling copy first), or None."""
    for d in (REF_DIR, SOURCE):
        if os.path.isfile(os.path.join(d, 'workflows', 'pseudo_label_generator.py')):
            return d
    return None


_installed = None


def install_shim(ref_dir=None):
    global _installed
    import torch
    ref_dir = ref_dir or available()
    if ref_dir is None:
        raise RuntimeError('the reference sources are not available (neither oracle/_ref/code nor /root/reference/code)')
    if _installed == ref_dir:
        return ref_dir
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_dir)
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
    if not os.path.isfile(os.path.join(ref_dir, 'utils', 'utils.py')):
        # the travelling copy holds the hot-path files only: utils/utils.py (model loading, logging, DDP helpers -- used by
        # BasePseudoGenerator.initialize, which the harness replaces) and the backbone builder are stand-ins
        import importlib
        importlib.import_module('utils')
        uu = ModuleType('utils.utils')
        uu.load_model = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('utils.load_model is outside oracle/_ref'))
        uu.create_dir = lambda d: os.makedirs(d, exist_ok=True)
        sys.modules['utils.utils'] = uu
        sys.modules['utils'].utils = uu
        sm = ModuleType('sseg.models.modules.seg_models')
        sm.build_seg_model = lambda cfg: None
        sys.modules['sseg.models.modules.seg_models'] = sm
    _installed = ref_dir
    return ref_dir


def run_ias(batches, C, alpha, beta, gamma, cp_gamma, keep_labels=True):
    """``IASPseudoGenerator.run`` of the unmodified reference on ``batches`` = [(logits f32 [B,C,H,W] on the CPU, paths)]."""
    install_shim()
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
            if keep_labels:
                self.captured.append(plbl.astype(np.uint8))   # the PNG payload of :46
            self.thr_per_image.append(self.class_threshold.copy())

        def save_data(self):
            pass

    cfg = SimpleNamespace(
        dataset=SimpleNamespace(num_classes=C),
        pseudo_policy=SimpleNamespace(type='IAS', ias=SimpleNamespace(alpha=alpha, beta=beta, gamma=gamma)),
        preprocessor=SimpleNamespace(copy_paste=SimpleNamespace(gamma=cp_gamma)))
    gen = Harness(cfg)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        gen.run()
    shutil.rmtree(gen.pseudo_label_save_dir, ignore_errors=True)
    return gen


if __name__ == '__main__':
    print(build(verbose=True))
