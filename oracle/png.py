"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the pseudo-label PNG writer.

The reference saves every pseudo-label with ``cv2.imwrite(path, plbl.astype(np.uint8))``
(code/workflows/pseudo_label_generator.py:43-46) and reads it back with ``np.array(Image.open(path))``
(code/sseg/datasets/loader/base_dataset.py:158-170).  The contract is therefore the DECODED image: a
single-channel 8-bit PNG whose pixels equal the label map.  libpng/zlib (what cv2 links) are not part of
/root/reference; PNG (ISO/IEC 15948) and DEFLATE / zlib (RFC 1951 / 1950) are published formats and any
conforming stream decodes to the same pixels.  Parity is pinned two ways in tests/test_png_host.py:
every stream made here decodes through cv2.imread, PIL and zlib to the input array, and cv2.imwrite's own
output for the same array decodes to the same pixels.

This module restates the CUDA encoder's (hiast_b200/csrc/png.cu) stream layout bit for bit, so the GPU
tests can compare FILE BYTES, not just decoded pixels:

* raw stream: per row one filter byte 2 (Up) + W bytes label[r] - label[r-1] mod 256 (row -1 = zeros): rows
  that repeat the row above become zeros, which is what makes segmentation maps compress;
* an image is cut into segments of R = 256 // ceil(W/128) rows; every segment is ONE IDAT chunk;
* a row is cut into chunks of 128 filtered bytes; a chunk is tokenised on its own: a byte equal to its left
  neighbour (the neighbour may be in the previous chunk, never in the previous row) extends a run, runs of
  >= 3 become one match (length = run, distance 1), everything else is a literal; the filter byte is a
  literal 2 in front of the row's first chunk;
* segment data, 'fixed' mode: block header (BFINAL 0, BTYPE 01), the tokens in the fixed Huffman code,
  end-of-block, then an empty stored block (sync flush: 000, pad, 00 00 FF FF) so that the next segment
  starts on a byte boundary; 'stored' mode (chosen when it is strictly smaller): one stored block with the
  raw bytes;
* the first segment carries the zlib header 78 01 in front, the last one a final empty fixed block (03 00)
  and the Adler-32 of the raw stream behind.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

from __future__ import annotations

import struct
import zlib

import numpy as np

CHUNK = 128
MAX_CHUNKS = 256
SIGNATURE = b'\x89PNG\r\n\x1a\n'
IEND = b'\x00\x00\x00\x00IEND\xaeB`\x82'


def geometry(H, W):
    """(chunks per row, rows per segment, segments per image)."""
    cpr = (W + CHUNK - 1) // CHUNK
    if cpr > MAX_CHUNKS or H < 1 or W < 1:
        raise ValueError('unsupported image size %dx%d' % (H, W))
    R = min(H, MAX_CHUNKS // cpr)
    return cpr, R, (H + R - 1) // R


def max_file_bytes(H, W):
    """Upper bound of the file size (every segment stored)."""
    _, _, S = geometry(H, W)
    return 8 + 25 + 12 + S * (12 + 5) + H * (W + 1) + 2 + 6


def _rev(code, n):
    r = 0
    for _ in range(n):
        r = (r << 1) | (code & 1)
        code >>= 1
    return r


_LIT = [(_rev(0x30 + b, 8), 8) if b < 144 else (_rev(0x190 + b - 144, 9), 9) for b in range(256)]


def _match_token(run):
    """(value, nbits) of a match of length `run` (3..258) at distance 1, RFC 1951 3.2.5 / 3.2.6."""
    if run == 258:
        sym, e, extra = 285, 0, 0
    else:
        l = run - 3
        if l < 8:
            sym, e, extra = 257 + l, 0, 0
        else:
            e = l.bit_length() - 1 - 2
            sym = 261 + 4 * e + ((l >> e) & 3)
            extra = l & ((1 << e) - 1)
    if sym < 280:
        code, nb = sym - 256, 7
    else:
        code, nb = 0xC0 + sym - 280, 8
    return _rev(code, nb) | (extra << nb), nb + e + 5            # + 5 zero bits: distance code 0


_MATCH = [None, None, None] + [_match_token(r) for r in range(3, 259)]


FILTER_UP = 2


def up_filter(label):
    """PNG filter type 2 of every row: label[r] - label[r-1] mod 256, row -1 = zeros (ISO/IEC 15948 9.2)."""
    f = label.copy()
    f[1:] -= label[:-1]
    return f


def _segment_tokens(seg):
    """Token (value, nbits) arrays of a segment [rows, W] of FILTERED bytes in stream order."""
    rows, W = seg.shape
    vals, nbs = [], []
    for r in range(rows):
        row = seg[r]
        vals.append(_LIT[FILTER_UP][0])
        nbs.append(8)
        for x0 in range(0, W, CHUNK):
            n = min(CHUNK, W - x0)
            i = 0
            while i < n:
                x = x0 + i
                b = int(row[x])
                if x > 0 and b == int(row[x - 1]):
                    run = 1
                    while i + run < n and int(row[x + run]) == b:
                        run += 1
                    if run >= 3:
                        v, nb = _MATCH[run]
                        vals.append(v)
                        nbs.append(nb)
                        i += run
                        continue
                vals.append(_LIT[b][0])
                nbs.append(_LIT[b][1])
                i += 1
    return np.asarray(vals, dtype=np.int64), np.asarray(nbs, dtype=np.int64)


def _pack_bits(vals, nbs, lead_bits):
    """LSB-first bit packing of the tokens behind `lead_bits` (list of 0/1); returns (bytes, total bits)."""
    off = np.concatenate([[0], np.cumsum(nbs)]) + len(lead_bits)
    total = int(off[-1])
    bits = np.zeros(total, dtype=np.uint8)
    bits[:len(lead_bits)] = lead_bits
    for k in range(int(nbs.max()) if len(nbs) else 0):
        m = nbs > k
        bits[off[:-1][m] + k] = (vals[m] >> k) & 1
    return bits, total


def _segment_data(seg):
    """DEFLATE bytes of one segment and its mode."""
    rows, W = seg.shape
    L = rows * (W + 1)
    vals, nbs = _segment_tokens(seg)
    bits, total = _pack_bits(vals, nbs, [0, 1, 0])
    total += 7 + 3                                              # end of block + empty stored block header
    nbytes = (total + 7) // 8
    fixed_len = nbytes + 4
    stored_len = 5 + L
    if fixed_len > stored_len:
        raw = np.full((rows, W + 1), FILTER_UP, dtype=np.uint8)
        raw[:, 1:] = seg
        return b'\x00' + struct.pack('<HH', L, L ^ 0xFFFF) + raw.tobytes(), 'stored'
    padded = np.zeros(nbytes * 8, dtype=np.uint8)
    padded[:len(bits)] = bits
    return np.packbits(padded, bitorder='little').tobytes() + b'\x00\x00\xff\xff', 'fixed'


def _chunk(kind, data):
    return struct.pack('>I', len(data)) + kind + data + struct.pack('>I', zlib.crc32(kind + data) & 0xFFFFFFFF)


def encode_png(label, return_modes=False):
    """PNG file bytes of a uint8 [H, W] label map, identical to hiast_png_encode's output."""
    label = np.ascontiguousarray(label, dtype=np.uint8)
    H, W = label.shape
    _, R, S = geometry(H, W)
    flt = up_filter(label)
    raw = np.full((H, W + 1), FILTER_UP, dtype=np.uint8)
    raw[:, 1:] = flt
    adler = zlib.adler32(raw.tobytes()) & 0xFFFFFFFF
    out = [SIGNATURE, _chunk(b'IHDR', struct.pack('>IIBBBBB', W, H, 8, 0, 0, 0, 0))]
    modes = []
    for s in range(S):
        data, mode = _segment_data(flt[s * R:min(H, (s + 1) * R)])
        modes.append(mode)
        if s == 0:
            data = b'\x78\x01' + data
        if s == S - 1:
            data = data + b'\x03\x00' + struct.pack('>I', adler)
        out.append(_chunk(b'IDAT', data))
    out.append(IEND)
    blob = b''.join(out)
    return (blob, modes) if return_modes else blob


def decode_png(blob):
    """Minimal PNG reader (gray 8-bit, filters None/Up) on top of zlib: one of the independent decoders of the tests."""
    assert blob[:8] == SIGNATURE
    pos, idat, W, H = 8, [], None, None
    while pos < len(blob):
        n, kind = struct.unpack('>I4s', blob[pos:pos + 8])
        data = blob[pos + 8:pos + 8 + n]
        crc, = struct.unpack('>I', blob[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(kind + data) & 0xFFFFFFFF, 'bad CRC in %r' % kind
        if kind == b'IHDR':
            W, H, depth, ctype, comp, flt, inter = struct.unpack('>IIBBBBB', data)
            assert (depth, ctype, comp, flt, inter) == (8, 0, 0, 0, 0)
        elif kind == b'IDAT':
            idat.append(data)
        elif kind == b'IEND':
            assert pos + 12 + n == len(blob)
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(b''.join(idat)), dtype=np.uint8).reshape(H, W + 1)
    img = raw[:, 1:].copy()
    for r in range(H):
        assert raw[r, 0] in (0, 2)
        if raw[r, 0] == 2 and r > 0:
            img[r] += img[r - 1]
    return img
