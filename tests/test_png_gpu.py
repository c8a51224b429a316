"""GPU parity of the device PNG writer (hiast_png_encode) against oracle/png.py: FILE BYTES identical, and the files
decode (cv2 / PIL, the reference's reader) to the label maps, incl. full-size 1024x2048 batches."""

import io

import numpy as np
import pytest
import torch

from oracle import png as opng
from test_png_host import SIZES, label_maps

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip('cv2')
Image = pytest.importorskip('PIL.Image')


def encode(maps, H, W, **kw):
    from hiast_b200.ops import PngEncoder
    enc = PngEncoder(H, W, max_images=len(maps), **kw)
    dev = torch.from_numpy(np.stack(maps)).cuda()
    return [bytes(f) for f in enc.encode_to_host(dev)], enc


@pytest.mark.parametrize('H,W', SIZES + [(21, 1536), (64, 2048), (2, 4000)])
def test_bytes_equal_oracle(H, W):
    maps = label_maps(H, W, H * 1000 + W)
    files, enc = encode(list(maps.values()), H, W)
    for (name, lbl), blob in zip(maps.items(), files):
        want = opng.encode_png(lbl)
        assert len(blob) == len(want), (name, len(blob), len(want))
        assert blob == want, name
        assert np.array_equal(np.array(Image.open(io.BytesIO(blob)), dtype=np.uint8), lbl)
    assert enc.max_file == opng.max_file_bytes(H, W) and enc.segments == opng.geometry(H, W)[2]


def test_unaligned_label_pointer_and_odd_batch():
    """A label view that is not 16-byte aligned takes the byte loader; result unchanged."""
    from hiast_b200.ops import PngEncoder
    H, W = 37, 256
    lbl = label_maps(H, W, 1)['mixed']
    buf = torch.zeros(H * W * 3 + 1, dtype=torch.uint8, device='cuda')
    view = buf[1:1 + 3 * H * W].view(3, H, W)
    view.copy_(torch.from_numpy(np.stack([lbl, lbl[::-1].copy(), lbl])).cuda())
    enc = PngEncoder(H, W, max_images=8)
    files = [bytes(f) for f in enc.encode_to_host(view)]
    assert files[0] == opng.encode_png(lbl) == files[2]
    assert files[1] == opng.encode_png(lbl[::-1].copy())


def test_small_capacity_grows():
    H, W = 64, 512
    noise = label_maps(H, W, 2)['noise']
    files, enc = encode([noise] * 4, H, W, expect_ratio=50.0)      # stored mode: needs the worst-case buffer
    assert all(f == opng.encode_png(noise) for f in files)


def test_full_size_batch_round_trip():
    """BASELINE size: 8 maps of 1024x2048 -- decoded pixels equal the maps (cv2 + PIL); one file byte-compared."""
    H, W = 1024, 2048
    g = torch.Generator(device='cuda').manual_seed(3)
    low = torch.randn(8, 19, 32, 64, generator=g, device='cuda') * 4
    logits = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True)
    logits += torch.randn(logits.shape, generator=g, device='cuda') * 0.5
    conf, lbl = torch.softmax(logits, 1).max(1)
    lbl = lbl.to(torch.uint8)
    lbl[conf < 0.9] = 255
    from hiast_b200.ops import PngEncoder
    enc = PngEncoder(H, W, max_images=8)
    files = [bytes(f) for f in enc.encode_to_host(lbl)]
    host = lbl.cpu().numpy()
    for i, blob in enumerate(files):
        a = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(a, host[i])
        assert len(blob) < H * W // 4
    assert np.array_equal(np.array(Image.open(io.BytesIO(files[3])), dtype=np.uint8), host[3])
    assert files[0] == opng.encode_png(host[0])


def test_invalid_arguments():
    from hiast_b200 import _lib
    from hiast_b200.ops import PngEncoder
    l = _lib.lib()
    assert l.hiast_png_encode(None, 1, 4, 4, None, 0, None, None, 0, None) == -1
    assert l.hiast_png_max_bytes(4, 128 * 256 + 1) == 0
    with pytest.raises(_lib.HiastError):
        PngEncoder(8, 8, 2).encode(torch.zeros(1, 8, 8, dtype=torch.uint8))     # CPU tensor: no host path
    with pytest.raises(_lib.HiastError):
        PngEncoder(8, 8, 2).encode(torch.zeros(3, 8, 8, dtype=torch.uint8, device='cuda'))


def test_random_shapes_and_contents_equal_oracle_bytes():
    """The same seeded sweep as the host test, device bytes against oracle bytes (runs of every length, every chunk tail)."""
    from hiast_b200.ops import PngEncoder
    rng = np.random.default_rng(2024)
    for trial in range(60):
        H = int(rng.integers(1, 70))
        W = int(rng.choice([1, 2, 3, 4, 5, 127, 128, 129, 130, 255, 256, 257, 383, 384, 385, 511, 513, int(rng.integers(1, 700))]))
        kind = trial % 4
        if kind == 0:
            flat = np.repeat(rng.integers(0, 256, H * W), rng.geometric(0.15, H * W))[:H * W]
        elif kind == 1:
            row = rng.integers(0, 19, W)
            flat = np.concatenate([np.where(rng.random(W) < 0.02 * (r % 5), rng.integers(0, 19, W), row) for r in range(H)])
        elif kind == 2:
            flat = np.where(rng.random(H * W) < 0.8, 7, 255)
        else:
            flat = rng.integers(0, 256, H * W)
        lbl = flat.astype(np.uint8).reshape(H, W)
        got = bytes(PngEncoder(H, W, 1).encode_to_host(torch.from_numpy(lbl).cuda())[0])
        assert got == opng.encode_png(lbl), (trial, H, W, kind)


def test_many_small_images_and_tall_images():
    """More than 256 images in one call (block-wide offset scan in several rounds) and more than 32 segments per image."""
    from hiast_b200.ops import PngEncoder
    rng = np.random.default_rng(8)
    maps = rng.integers(0, 4, (300, 6, 10)).astype(np.uint8)
    files = PngEncoder(6, 10, 300).encode_to_host(torch.from_numpy(maps).cuda())
    assert len(files) == 300
    for i in (0, 1, 255, 256, 257, 299):
        assert bytes(files[i]) == opng.encode_png(maps[i]), i
    tall = np.repeat(rng.integers(0, 19, (2, 90, 30)), 100, 2).astype(np.uint8)       # [2, 90, 3000]: 24 chunks/row, R = 10, S = 9
    wide = np.repeat(rng.integers(0, 19, (1, 700, 40)), 100, 2).astype(np.uint8)      # [1, 700, 4000]: 32 chunks/row, R = 8, S = 88
    for arr in (tall, wide):
        n, H, W = arr.shape
        files = PngEncoder(H, W, n).encode_to_host(torch.from_numpy(arr).cuda())
        for i in range(n):
            assert bytes(files[i]) == opng.encode_png(arr[i])
