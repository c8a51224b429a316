"""CPU tests of the PNG oracle (oracle/png.py): every stream decodes to the input through cv2, PIL (the reference's
reader, base_dataset.py:158-170) and a zlib-based reader, and agrees with what cv2.imwrite (the reference's writer,
pseudo_label_generator.py:43-46) stores for the same array."""

import io
import os
import zlib

import numpy as np
import pytest

from oracle import png

cv2 = pytest.importorskip('cv2')
Image = pytest.importorskip('PIL.Image')


def label_maps(H, W, seed):
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 19, (H, W)).astype(np.uint8)
    noise[rng.random((H, W)) < 0.5] = 255
    blocks = np.repeat(np.repeat(rng.integers(0, 19, ((H + 7) // 8, (W + 15) // 16)), 8, 0), 16, 1)[:H, :W].astype(np.uint8)
    mixed = blocks.copy()
    m = rng.random((H, W)) < 0.03
    mixed[m] = 255
    const = np.full((H, W), 255, np.uint8)
    ramp = (np.arange(H * W) % 251).astype(np.uint8).reshape(H, W)
    return {'noise': noise, 'blocks': blocks, 'mixed': mixed, 'const': const, 'ramp': ramp}


SIZES = [(1, 1), (1, 2), (2, 3), (3, 5), (7, 127), (5, 128), (4, 129), (16, 256), (17, 300), (40, 1000), (33, 2048), (300, 64)]


@pytest.mark.parametrize('H,W', SIZES)
def test_oracle_stream_decodes_everywhere(H, W):
    for name, lbl in label_maps(H, W, H * 1000 + W).items():
        blob = png.encode_png(lbl)
        assert len(blob) <= png.max_file_bytes(H, W)
        assert np.array_equal(png.decode_png(blob), lbl), name
        a = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        assert a is not None and a.dtype == np.uint8 and np.array_equal(a.reshape(H, W), lbl), name
        b = np.array(Image.open(io.BytesIO(blob)), dtype=np.uint8)        # base_dataset.py:166
        assert np.array_equal(b, lbl), name


def test_same_pixels_as_the_reference_writer(tmp_path):
    """cv2.imwrite -> Image.open (the reference's round trip) and oracle -> Image.open give the same array."""
    lbl = label_maps(96, 200, 5)['mixed']
    ref_path = os.path.join(tmp_path, 'ref_pseudo_label.png')
    cv2.imwrite(ref_path, lbl.astype(np.uint8))                            # pseudo_label_generator.py:46
    ours_path = os.path.join(tmp_path, 'ours_pseudo_label.png')
    with open(ours_path, 'wb') as f:
        f.write(png.encode_png(lbl))
    a = np.array(Image.open(ref_path), dtype=np.uint8)
    b = np.array(Image.open(ours_path), dtype=np.uint8)
    assert a.shape == b.shape and np.array_equal(a, b) and np.array_equal(a, lbl)
    assert Image.open(ours_path).mode == Image.open(ref_path).mode == 'L'


def test_modes_and_sizes():
    maps = label_maps(64, 512, 3)
    _, modes = png.encode_png(maps['noise'], return_modes=True)
    assert set(modes) == {'stored'}                                        # Up-filtered noise does not compress
    blob, modes = png.encode_png(maps['blocks'], return_modes=True)
    assert set(modes) == {'fixed'} and len(blob) < 64 * 512 // 10
    blob, modes = png.encode_png(maps['const'], return_modes=True)
    assert set(modes) == {'fixed'} and len(blob) < 64 * 512 // 40
    cpr, R, S = png.geometry(1024, 2048)
    assert (cpr, R, S) == (16, 16, 64)
    assert png.geometry(768, 1536) == (12, 21, 37)
    with pytest.raises(ValueError):
        png.geometry(4, 128 * 256 + 1)


def test_golden_stream_hash():
    """The byte layout is frozen: a change of the tokeniser / framing shows up here and in the GPU byte comparison."""
    lbl = label_maps(48, 300, 11)['mixed']
    blob = png.encode_png(lbl)
    assert (len(blob), zlib.crc32(blob)) == GOLDEN


GOLDEN = (3505, 1597620267)


def test_random_shapes_and_contents_round_trip():
    """Seeded sweep over sizes around the chunk / segment boundaries and contents with runs of every length."""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        H = int(rng.integers(1, 70))
        W = int(rng.choice([1, 2, 3, 4, 5, 127, 128, 129, 130, 255, 256, 257, 383, 384, 385, 511, 513, int(rng.integers(1, 700))]))
        kind = trial % 4
        if kind == 0:                                   # runs of random length (geometric), random values
            flat = np.repeat(rng.integers(0, 256, H * W), rng.geometric(0.15, H * W))[:H * W]
        elif kind == 1:                                 # vertical structure: rows repeat with sparse changes
            row = rng.integers(0, 19, W)
            flat = np.concatenate([np.where(rng.random(W) < 0.02 * (r % 5), rng.integers(0, 19, W), row) for r in range(H)])
        elif kind == 2:                                 # two-valued noise (matches and literals interleave at every length)
            flat = np.where(rng.random(H * W) < 0.8, 7, 255)
        else:
            flat = rng.integers(0, 256, H * W)
        lbl = flat.astype(np.uint8).reshape(H, W)
        blob = png.encode_png(lbl)
        assert np.array_equal(png.decode_png(blob), lbl), (H, W, kind)
        a = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(a.reshape(H, W), lbl), (H, W, kind)
