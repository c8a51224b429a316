"""Validator path (workflows/validator.py:34-55,78-115): oracle vs the reference fixture on CPU; fused kernels vs the
reference's own CUDA path (torch ops via the oracle) bit for bit on the GPU."""

import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import metrics as omet
from oracle import validator as oval

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'validator.npz')


@pytest.mark.parametrize('key', list(gi.VALIDATOR_SPECS))
def test_oracle_equals_reference_fixture(key):
    g = np.load(GOLD)
    spec = gi.VALIDATOR_SPECS[key]
    model = gi.ToyModel(spec['C'], spec['seed'])
    imgs = gi.validator_images(spec)
    with torch.no_grad():
        res = oval.multi_scale_and_flip(model, imgs, spec['sizes'], spec['flip'])
    assert np.array_equal(res.numpy(), g[key + '_results'])
    assert np.array_equal(res.argmax(1).numpy().astype(np.uint8), g[key + '_labels'])


def cfg_of(spec, color_dir=None, source='GTA5'):
    return SimpleNamespace(dataset=SimpleNamespace(num_classes=spec['C'], source=SimpleNamespace(type=source)),
                           validate=SimpleNamespace(resize_sizes=spec['sizes'], is_flip=spec['flip'], batch_size=spec['B'],
                                                    color_mask_dir_path=color_dir))


@pytest.mark.gpu
@pytest.mark.parametrize('key', list(gi.VALIDATOR_SPECS))
def test_fused_prediction_equals_torch_cuda_path(key):
    from hiast_b200.validator import Validator
    spec = gi.VALIDATOR_SPECS[key]
    model = gi.ToyModel(spec['C'], spec['seed'])
    imgs = gi.validator_images(spec).cuda()
    v = Validator(cfg_of(spec), model=model, loader=[])
    with torch.no_grad():
        want = oval.multi_scale_and_flip(model, imgs, spec['sizes'], spec['flip'])
        got_sum = v.get_multi_scale_and_flip_logits(imgs)
        got = v.predict(imgs)
    assert torch.equal(got_sum, want)                              # softmax(+flip) kernel bit-exact vs ATen
    assert torch.equal(got.long(), want.argmax(1))
    # the CPU fixture of the unmodified reference agrees except where CPU and CUDA softmax differ in the last ulp
    g = np.load(GOLD)
    assert (got.cpu().numpy() != g[key + '_labels']).mean() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize('B,C,H,W,sizes,flip', [
    (2, 19, 128, 256, [[96, 192]], False),                         # configs/validate.yaml shape family (768x1536 -> 1024x2048)
    (1, 19, 128, 256, [[96, 192], [128, 256], [160, 320]], True),
    (1, 19, 1024, 2048, [[768, 1536]], True),                      # full size
    (2, 9, 65, 131, [[33, 67], [65, 131]], True),                  # odd widths: scalar path
    (1, 3, 8, 8, [[8, 8]], False),
])
def test_kernels_equal_aten_on_random_logits(B, C, H, W, sizes, flip):
    from hiast_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(H * 7 + W)
    from torch.nn import functional as F
    probs, want = [], 0
    for (h, w) in sizes:
        z0 = torch.randn(B, C, h, w, generator=g, device='cuda') * 4
        z1 = torch.randn(B, C, h, w, generator=g, device='cuda') * 4 if flip else None
        ref = F.softmax(z0, dim=1)
        if flip:
            ref += torch.flip(F.softmax(z1, dim=1), dims=[3])
        got = ops.softmax_flip_sum(z0, z1)
        assert torch.equal(got, ref)
        probs.append(got)
        want = want + F.interpolate(ref, (H, W), mode='bilinear', align_corners=True)
    lbl = ops.probs_upsample_argmax(probs, (H, W))                 # staged kernel (<= 3 scales)
    assert torch.equal(lbl.long(), want.argmax(1))
    from hiast_b200._lib import lib
    lib().hiast_debug_validate_direct(1)
    try:
        lbl_direct = ops.probs_upsample_argmax(probs, (H, W))      # direct kernel
    finally:
        lib().hiast_debug_validate_direct(0)
    assert torch.equal(lbl_direct, lbl)


@pytest.mark.gpu
def test_five_scales_and_downsampling_take_the_general_paths():
    from hiast_b200 import ops
    from torch.nn import functional as F
    g = torch.Generator(device='cuda').manual_seed(4)
    B, C, H, W = 2, 6, 40, 72
    sizes = [(20, 36), (40, 72), (64, 128), (30, 50), (80, 144)]         # up- and down-sampling, widths not multiples of 4
    probs = [torch.softmax(torch.randn(B, C, h, w, generator=g, device='cuda') * 3, 1) for h, w in sizes]
    for k in (5, 3, 1):
        want = sum(F.interpolate(p, (H, W), mode='bilinear', align_corners=True) for p in probs[:k])
        assert torch.equal(ops.probs_upsample_argmax(probs[:k], (H, W)).long(), want.argmax(1)), k
    big = [torch.softmax(torch.randn(1, 3, 512, 1024, generator=g, device='cuda'), 1)]   # 8x down-sampling: wide windows
    want = F.interpolate(big[0], (64, 128), mode='bilinear', align_corners=True)
    assert torch.equal(ops.probs_upsample_argmax(big, (64, 128)).long(), want.argmax(1))


@pytest.mark.gpu
def test_argmax_ties_take_the_first_index():
    from hiast_b200 import ops
    p = torch.zeros(1, 5, 6, 8, device='cuda')
    p[:, 1] = 0.5
    p[:, 3] = 0.5
    p[:, 4, :, 4:] = 0.5
    p[:, 0, 2, :] = 0.5
    lbl = ops.probs_upsample_argmax([p], (6, 8))
    assert torch.equal(lbl.long(), p.argmax(1))
    assert lbl[0, 0, 0] == 1 and lbl[0, 2, 0] == 0


@pytest.mark.gpu
def test_validator_run_matches_oracle_miou(tmp_path):
    from hiast_b200.validator import Validator
    spec = dict(gi.VALIDATOR_SPECS['multi_scale_flip'])
    model = gi.ToyModel(spec['C'], 5)
    g = torch.Generator().manual_seed(77)
    batches = []
    for i in range(3):
        imgs = torch.randn(2, 3, spec['H'], spec['W'], generator=g)
        lbls = torch.randint(0, spec['C'], (2, spec['H'], spec['W']), generator=g)
        lbls[torch.rand(lbls.shape, generator=g) < 0.1] = 255
        batches.append({'images': imgs, 'labels': lbls, 'image_paths': ['/x/img_%d_%d.png' % (i, k) for k in range(2)]})
    color_dir = str(tmp_path / 'color')
    v = Validator(cfg_of(spec, color_dir, source='SYNTHIA'), model=model, loader=batches)
    res = v.run()
    inter = np.zeros(spec['C'], np.float32)
    union = np.zeros(spec['C'], np.float32)
    cm = np.zeros((spec['C'] + 1, spec['C'] + 1), np.int64)
    for b in batches:
        with torch.no_grad():
            pred = oval.predict_labels(model, b['images'].cuda(), spec['sizes'], spec['flip']).cpu().numpy()
        cm += omet.confusion_matrix(pred, b['labels'].numpy(), spec['C'])
        i_, u_, _ = omet.intersection_and_union(pred, b['labels'].numpy(), spec['C'])      # metrics.py:6-19
        inter += i_                                                                          # validator.py:96-97
        union += u_
    assert np.array_equal(res['confusion_matrix'].cpu().numpy(), cm)
    want = omet.iou_from_sums(inter, union, synthia=True)
    assert np.array_equal(res['iou'], want['iou'])
    assert res['miou_16'] == want['miou_16'] and res['miou_13'] == want['miou_13']
    assert sorted(os.listdir(color_dir)) == sorted('img_%d_%d.png' % (i, k) for i in range(3) for k in range(2))
    from PIL import Image
    assert Image.open(os.path.join(color_dir, 'img_0_0.png')).mode == 'P'
