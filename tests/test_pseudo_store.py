"""Reader side of the on-disk outputs (base_dataset.py:61-77,158-178): oracle and host mirror against the reference
fixture (CPU), device nearest resize against the fixture and cv2 (GPU), and the write -> read round trip."""

import json
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import pseudo_store as ops_oracle

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'pseudo_store.npz')
cv2 = pytest.importorskip('cv2')


def gold():
    return np.load(GOLD)


def test_oracle_stat_samples_equals_reference():
    want = {int(k): v for k, v in json.loads(str(gold()['stat_json'])).items()}
    got = ops_oracle.stat_samples_with_class(json.loads(json.dumps(gi.pseudo_store_samples())), gi.PSEUDO_STORE_SPEC['C'])
    assert got == want
    assert got[1] == [] and len(got[2]) == 4 and len(got[0]) == 23 - round(2.3)


def test_host_mirror_stat_samples_equals_reference(tmp_path):
    from hiast_b200 import pseudo_store
    with open(tmp_path / 'samples_with_class.json', 'a') as f:           # save_data opens in append mode (:61)
        f.write(json.dumps(gi.pseudo_store_samples()))
    want = {int(k): v for k, v in json.loads(str(gold()['stat_json'])).items()}
    assert pseudo_store.stat_samples_with_class(str(tmp_path), gi.PSEUDO_STORE_SPEC['C']) == want
    assert pseudo_store.pseudo_label_path('/p', '/a/b/frankfurt_000.png') == '/p/frankfurt_000_pseudo_label.png'


def test_oracle_nearest_resize_equals_reference_and_cv2():
    g = gold()
    for k, (src, dst) in enumerate(gi.PSEUDO_STORE_SPEC['sizes']):
        lbl = gi.pseudo_store_label(k, src)
        got = ops_oracle.resize_nearest(lbl, dst)
        assert np.array_equal(got, g['lbl_%d' % k]), (src, dst)
        assert np.array_equal(got, cv2.resize(lbl, dst[::-1], interpolation=cv2.INTER_NEAREST))


@pytest.mark.gpu
def test_device_nearest_resize_equals_reference_fixture():
    from hiast_b200 import ops
    g = gold()
    for k, (src, dst) in enumerate(gi.PSEUDO_STORE_SPEC['sizes']):
        lbl = gi.pseudo_store_label(k, src)
        got = ops.resize_nearest_u8(torch.from_numpy(lbl).cuda(), dst).cpu().numpy()
        assert np.array_equal(got, g['lbl_%d' % k]), (src, dst)


@pytest.mark.gpu
@pytest.mark.parametrize('src,dst', [((768, 1536), (1024, 2048)), ((1024, 2048), (512, 1024)), ((513, 1025), (1024, 2048)),
                                      ((760, 1280), (1024, 2047))])
def test_device_nearest_resize_full_size_equals_cv2(src, dst):
    from hiast_b200 import ops
    rng = np.random.default_rng(5)
    lbl = rng.integers(0, 256, (3,) + src).astype(np.uint8)
    got = ops.resize_nearest_u8(torch.from_numpy(lbl).cuda(), dst).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], cv2.resize(lbl[i], dst[::-1], interpolation=cv2.INTER_NEAREST))
        assert np.array_equal(got[i], ops_oracle.resize_nearest(lbl[i], dst))


@pytest.mark.gpu
def test_write_then_read_round_trip(tmp_path):
    """Files written by the device PNG encoder come back through load_pseudo_labels (PIL + device resize) as the
    reference's load_data would return them."""
    from hiast_b200 import ops, pseudo_store
    rng = np.random.default_rng(9)
    H, W = 96, 192
    lbl = np.repeat(np.repeat(rng.integers(0, 19, (5, H // 8, W // 8)), 8, 1), 8, 2).astype(np.uint8)
    lbl[rng.random(lbl.shape) < 0.05] = 255
    files = ops.PngEncoder(H, W, 5).encode_to_host(torch.from_numpy(lbl).cuda())
    paths = ['/data/x/img_%d.png' % i for i in range(5)]
    for p, f in zip(paths, files):
        with open(pseudo_store.pseudo_label_path(str(tmp_path), p), 'wb') as fh:
            fh.write(bytes(f))
    same = pseudo_store.load_pseudo_labels(str(tmp_path), paths, (H, W))
    assert np.array_equal(same.cpu().numpy(), lbl)
    big = pseudo_store.load_pseudo_labels(str(tmp_path), paths, (128, 256)).cpu().numpy()
    for i in range(5):
        assert np.array_equal(big[i], cv2.resize(lbl[i], (256, 128), interpolation=cv2.INTER_NEAREST))
