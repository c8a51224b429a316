"""Seeded synthetic inputs shared by tests/golden/make_golden.py and the tests.

Nothing here touches the reference or the oracle.  All randomness comes from CPU
``torch.Generator`` / ``np.random.RandomState`` objects seeded from the spec, so the
GPU box regenerates exactly what the build container fed to the reference (the
fixtures store a sha256 of each input to detect generator drift).
"""

from __future__ import annotations

import numpy as np
import torch
from torch.nn import functional as F

IAS_SPECS = {
    # mixed diffuse / peaked maps, batch 2 with a trailing batch of 1 (2975 = 1487*2 + 1)
    'ias_small': dict(C=19, H=48, W=96, N=7, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99,
                      seed=11, dist='mixed', absent=()),
    # odd plane size (scalar tail paths), batch 3, default alpha/gamma of get_ias_threshold's
    # caller defaults (default_config.py:87-90: 0.2 / 0.9 / 8.0 -> here gamma 1.0), one class never predicted
    'ias_c7': dict(C=7, H=31, W=51, N=5, B=3, alpha=0.2, beta=0.9, gamma=1.0, cp_gamma=0.99,
                   seed=5, dist='peaked', absent=(5,)),
    # non-integer gamma exercises the general pow path
    'ias_g25': dict(C=19, H=32, W=64, N=4, B=2, alpha=0.35, beta=0.8, gamma=2.5, cp_gamma=0.9,
                    seed=23, dist='diffuse', absent=(9, 14, 16)),
    # BASELINE.json configs[0]: 8 x 19x512x1024, batch 2, randn*3 from Generator(0)
    'ias_config0': dict(C=19, H=512, W=1024, N=8, B=2, alpha=0.5, beta=0.9, gamma=8.0, cp_gamma=0.99,
                        seed=0, dist='baseline', absent=(), store_conf=False),
}

LOSS_SPECS = {
    'loss_ignored': dict(B=2, C=19, H=24, W=40, seed=3, p_ignore=0.5, region='ignored',
                         w_seg=1.0, w_kld=0.1, w_ent=1.0, w_cst=0.5, source=False),
    'loss_confident_src': dict(B=1, C=19, H=17, W=23, seed=4, p_ignore=0.3, region='confident',
                               w_seg=1.0, w_kld=0.1, w_ent=3.0, w_cst=0.5, source=True),
    'loss_all_c7': dict(B=3, C=7, H=8, W=20, seed=6, p_ignore=0.7, region='all',
                        w_seg=0.5, w_kld=0.2, w_ent=1.0, w_cst=1.0, source=False),
}

CST_VARIANT_SPEC = dict(B=2, C=19, H=12, W=20, seed=12, p_ignore=0.5, source=False)

METRIC_SPECS = {
    'metric_k19': dict(B=2, H=64, W=96, K=19, seed=7, p_ignore=0.1, p_oor=0.0),
    'metric_k16_oor': dict(B=1, H=33, W=57, K=16, seed=8, p_ignore=0.2, p_oor=0.05),
}

COPY_PASTE_SPEC = dict(C=19, H=32, W=48, n_img=6, n_run=4, selected=14, seed=9)


# ----------------------------------------------------------------------------- IAS
def diffuse_logits(g, n, C, H, W, scale=3.0):
    return torch.randn(n, C, H, W, generator=g) * scale


def peaked_logits(g, n, C, H, W, down=8, scale=4.0, noise=0.5):
    """Spatially coherent classes: low-res noise upsampled (align_corners=True) + fine noise."""
    h, w = max(H // down, 2), max(W // down, 2)
    low = torch.randn(n, C, h, w, generator=g) * scale
    up = F.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)
    return up + torch.randn(n, C, H, W, generator=g) * noise


def ias_batches(spec):
    """list of (logits f32 [b,C,H,W] CPU tensor, [paths]) in processing order."""
    g = torch.Generator().manual_seed(spec['seed'])
    C, H, W, N, B = spec['C'], spec['H'], spec['W'], spec['N'], spec['B']
    out = []
    i = 0
    k = 0
    while i < N:
        b = min(B, N - i)
        dist = spec['dist']
        if dist == 'baseline':
            lg = torch.randn(b, C, H, W, generator=g) * 3
        elif dist == 'diffuse' or (dist == 'mixed' and k % 2 == 0):
            lg = diffuse_logits(g, b, C, H, W)
        else:
            lg = peaked_logits(g, b, C, H, W)
        for c in spec['absent']:
            lg[:, c] = -1e4
        out.append((lg.contiguous(), ['img_%05d.png' % (i + j) for j in range(b)]))
        i += b
        k += 1
    return out


# ---------------------------------------------------------------------------- loss
def loss_inputs(spec):
    g = torch.Generator().manual_seed(spec['seed'])
    B, C, H, W = spec['B'], spec['C'], spec['H'], spec['W']
    z = torch.randn(B, C, H, W, generator=g) * 3
    t = torch.softmax(torch.randn(B, C, H, W, generator=g) * 3, dim=1)
    plbl = torch.randint(0, C, (B, H, W), generator=g)
    plbl[torch.rand(B, H, W, generator=g) < spec['p_ignore']] = 255
    s_z = s_lbl = None
    if spec['source']:
        s_z = torch.randn(B, C, H, W, generator=g) * 2
        s_lbl = torch.randint(0, C, (B, H, W), generator=g)
        s_lbl[torch.rand(B, H, W, generator=g) < 0.1] = 255
    return z, t, plbl, s_z, s_lbl


# -------------------------------------------------------------------------- metric
def metric_inputs(spec):
    g = torch.Generator().manual_seed(spec['seed'])
    B, H, W, K = spec['B'], spec['H'], spec['W'], spec['K']
    target = torch.randint(0, K, (B, H, W), generator=g)
    pred = target.clone()
    flip = torch.rand(B, H, W, generator=g) < 0.4
    pred[flip] = torch.randint(0, K, (int(flip.sum()),), generator=g)
    target[torch.rand(B, H, W, generator=g) < spec['p_ignore']] = 255
    if spec['p_oor'] > 0:
        pred[torch.rand(B, H, W, generator=g) < spec['p_oor']] = 255     # e.g. a pseudo-label map as pred
        target[torch.rand(B, H, W, generator=g) < spec['p_oor']] = K + 3  # label id outside [0,K)
    return pred, target


# ---------------------------------------------------------------------- copy-paste
def copy_paste_class_value(spec):
    rs = np.random.RandomState(spec['seed'] + 100)
    return rs.uniform(0.55, 0.99, size=spec['C'])


class CopyPasteDataset:
    """The three methods CopyPaste needs from its donor dataset (preprocessor.py:26,96-97)."""

    def __init__(self, spec):
        rs = np.random.RandomState(spec['seed'])
        C, H, W, n = spec['C'], spec['H'], spec['W'], spec['n_img']
        self.names = ['donor_%02d.png' % i for i in range(n)]
        self.imgs = [rs.randint(0, 256, size=(H, W, 3)).astype(np.uint8) for _ in range(n)]
        self.lbls = []
        for _ in range(n):
            coarse = rs.randint(0, C, size=(H // 4, W // 4))
            lbl = np.kron(coarse, np.ones((4, 4), dtype=np.int64)).astype(np.uint8)
            lbl[rs.rand(H, W) < 0.3] = 255
            self.lbls.append(lbl)
        self.samples = {c: [self.names[i] for i in range(n) if (self.lbls[i] == c).any()] or [self.names[0]]
                        for c in range(C)}

    def get_samples_with_class(self):
        return self.samples

    def get_file_to_idx(self, name):
        return self.names.index(name)

    def load_data(self, idx):
        return self.imgs[idx].copy(), self.lbls[idx].copy(), self.names[idx]


# ------------------------------------------------------------------ reader side of the on-disk outputs
PSEUDO_STORE_SPEC = dict(seed=21, C=5, sizes=[((18, 30), (24, 40)), ((12, 24), (16, 32)), ((7, 11), (24, 40)),
                                               ((30, 50), (24, 40)), ((24, 40), (24, 40)), ((96, 192), (128, 256)),
                                               ((1, 1), (5, 7)), ((33, 65), (100, 131))])


def pseudo_store_samples():
    """samples_with_class as save_data serialises it (pseudo_label_generator.py:60-62): {class: [[path, pixels], ...]},
    with ties, an empty class and a class with fewer than 5 files (round(len * 0.1) == 0)."""
    rng = np.random.default_rng(PSEUDO_STORE_SPEC['seed'])
    out = {}
    for c, n in enumerate([23, 0, 4, 15, 10]):
        px = rng.integers(1, 40, n)
        out[c] = [['/data/cityscapes/leftImg8bit/train/x/img_%d_%d.png' % (c, i), int(px[i])] for i in range(n)]
    return out


def pseudo_store_label(k, shape):
    rng = np.random.default_rng(100 + k)
    lbl = rng.integers(0, 19, shape).astype(np.uint8)
    lbl[rng.random(shape) < 0.3] = 255
    return lbl


# ------------------------------------------------------------------ validator (multi-scale / flip prediction)
VALIDATOR_SPECS = {
    'single_scale': dict(seed=31, B=2, C=19, H=16, W=32, sizes=[[12, 24]], flip=False),
    'multi_scale_flip': dict(seed=32, B=2, C=19, H=16, W=32, sizes=[[12, 24], [16, 32], [20, 40]], flip=True),
    'flip_c7_odd': dict(seed=33, B=1, C=7, H=15, W=33, sizes=[[9, 21], [15, 33]], flip=True),
}


class ToyModel:
    """Deterministic stand-in for the segmentation network: a fixed 1x1 projection of the image plus a position term,
    evaluated at the size of its input (what SelfTrainingSegmentor.forward returns after its own interpolate)."""

    def __init__(self, C, seed):
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.randn(C, 3, generator=g) * 2.0
        self.C = C

    def eval(self):
        return self

    def __call__(self, x):
        w = self.weight.to(x.device)
        b, _, h, wd = x.shape
        logits = torch.einsum('kc,bchw->bkhw', w, x)
        yy = torch.arange(h, device=x.device, dtype=torch.float32).view(1, 1, h, 1)
        xx = torch.arange(wd, device=x.device, dtype=torch.float32).view(1, 1, 1, wd)
        kk = torch.arange(self.C, device=x.device, dtype=torch.float32).view(1, self.C, 1, 1)
        logits = logits + torch.sin(0.37 * yy + 0.11 * xx * (kk + 1.0)) * 1.5
        return {'logits': logits.contiguous()}


def validator_images(spec):
    g = torch.Generator().manual_seed(spec['seed'])
    return torch.randn(spec['B'], 3, spec['H'], spec['W'], generator=g)


# ------------------------------------------------------------------ CE with class weights / refer_labels
CE_GENERAL_SPECS = {
    'b2_c19': dict(B=2, C=19, H=12, W=20, seed=41),
    'b3_c7': dict(B=3, C=7, H=9, W=11, seed=42),
    'b1_c19': dict(B=1, C=19, H=8, W=16, seed=43),
}


def ce_general_inputs(spec):
    g = torch.Generator().manual_seed(spec['seed'])
    B, C, H, W = spec['B'], spec['C'], spec['H'], spec['W']
    z = torch.randn(B, C, H, W, generator=g) * 3
    labels = torch.randint(0, C, (B, H, W), generator=g)
    weights = torch.rand(C, generator=g) * 2 + 0.1
    weights[1] = 0.0                                           # a zero-weight class: its pixels drop out of the non-zero count
    refer = torch.randint(0, C, (B, H, W), generator=g)
    refer[torch.rand(B, H, W, generator=g) < 0.5] = 255
    return z, labels, weights, refer


def ce_general_cases(labels, weights, refer):
    """keyword sets for LOSS['CE'](logits, **kw)."""
    with_ignore = labels.clone()
    with_ignore[refer == 255] = 255
    return {
        'weighted': dict(labels=with_ignore, weights=weights),
        'plain_ignore': dict(labels=with_ignore),
        'refer_ignored': dict(labels=labels, refer_labels=refer, region='ignored'),
        'refer_confident_w': dict(labels=labels, weights=weights, refer_labels=refer, region='confident'),
        'refer_all': dict(labels=labels, refer_labels=refer, region='all'),
    }
