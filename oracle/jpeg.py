"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of baseline-JPEG decoding as the reference performs it: `cv2.imread` at src/get_prob_video.py:95 reads the
face crops that src/data/get_face_images.py:60 wrote with `cv2.imwrite` defaults (baseline sequential DCT, 8 bit, YCbCr
4:2:0, the Annex-K Huffman tables, quality 95, no restart markers).  The arithmetic lives in a third-party dependency that
is not under /root/reference: libjpeg-turbo, bundled with opencv-python (pinned opencv-python==4.9.0.80 in
src/requirements.txt; 4.13.0 with libjpeg-turbo 3.1.2 in this image).  Its published algorithm, with the decompressor
defaults OpenCV leaves in place, is restated here:

  * entropy decoding          -- ITU T.81 F.2.2 (jdhuff.c): DC differences, AC run/size pairs, EOB / ZRL, EXTEND
  * dequantisation + IDCT     -- jidctint.c jpeg_idct_islow (JDCT_ISLOW, the default): 13-bit fixed-point Loeffler
                                 IDCT, two passes, PASS1_BITS = 2, post-IDCT range-limit table (wraps modulo 1024)
  * chroma upsampling         -- jdsample.c h2v2_fancy_upsample (do_fancy_upsampling = TRUE, the default): triangle
                                 filter, 3/4 - 1/4 taps with the +8 / +7 rounding alternation; the row above the first /
                                 below the last real chroma row is that row itself (jdmainct.c context rows)
  * colour conversion         -- jdcolor.c ycc_rgb_convert: 16-bit fixed-point tables, range limiting

The reference has no test for this; the restatement is pinned against `cv2.imdecode` itself (the reference's own decoder)
in tests/test_oracle_golden.py::test_jpeg_oracle_matches_cv2 on sizes that exercise partial MCUs and odd chroma widths.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], dtype=np.int64)          # zig-zag position -> natural (row-major) index


class JpegHeader:
    """What the markers say: frame size, per-component sampling / table selectors, quantisation and Huffman tables, and
    the entropy-coded segment (still byte-stuffed)."""

    def __init__(self):
        self.width = self.height = 0
        self.components: List[Tuple[int, int, int, int]] = []      # (id, h, v, tq)
        self.qt: Dict[int, np.ndarray] = {}                         # natural order, int32 [64]
        self.huff: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray]] = {}     # (class, id) -> (bits[17], values)
        self.scan: List[Tuple[int, int, int]] = []                  # (component index, dc table, ac table)
        self.restart_interval = 0
        self.data = b""


def parse(buf: bytes) -> JpegHeader:
    """Marker segments of a baseline file (T.81 B.2).  Raises ValueError for anything the path does not cover."""
    h = JpegHeader()
    if buf[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG (no SOI)")
    i = 2
    while i < len(buf):
        if buf[i] != 0xFF:
            raise ValueError(f"marker expected at byte {i}")
        m = buf[i + 1]
        if m == 0xFF:                      # fill byte
            i += 1
            continue
        if m == 0xD9:
            break
        seg_len = (buf[i + 2] << 8) | buf[i + 3]
        seg = buf[i + 4: i + 2 + seg_len]
        if m == 0xDB:                      # DQT
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                if pq:
                    raise ValueError("16-bit quantisation tables are not baseline")
                t = np.zeros(64, dtype=np.int32)
                t[ZIGZAG] = np.frombuffer(seg[j + 1: j + 65], dtype=np.uint8)
                h.qt[tq] = t
                j += 65
        elif m == 0xC0:                    # SOF0
            if seg[0] != 8:
                raise ValueError("only 8-bit samples")
            h.height, h.width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            for c in range(seg[5]):
                cid, hv, tq = seg[6 + 3 * c: 9 + 3 * c]
                h.components.append((cid, hv >> 4, hv & 15, tq))
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError(f"SOF marker 0x{m:02x}: only baseline sequential DCT (SOF0) is covered")
        elif m == 0xC4:                    # DHT
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                bits = np.zeros(17, dtype=np.int64)
                bits[1:] = np.frombuffer(seg[j + 1: j + 17], dtype=np.uint8)
                n = int(bits.sum())
                h.huff[(tc, th)] = (bits, np.frombuffer(seg[j + 17: j + 17 + n], dtype=np.uint8).astype(np.int64))
                j += 17 + n
        elif m == 0xDD:
            h.restart_interval = (seg[0] << 8) | seg[1]
        elif m == 0xDA:                    # SOS: the entropy-coded data follows up to EOI
            ns = seg[0]
            ids = [c[0] for c in h.components]
            for k in range(ns):
                cs, tt = seg[1 + 2 * k], seg[2 + 2 * k]
                h.scan.append((ids.index(cs), tt >> 4, tt & 15))
            end = buf.rfind(b"\xff\xd9")
            h.data = buf[i + 2 + seg_len: end if end > 0 else len(buf)]
            return h
        i += 2 + seg_len
    raise ValueError("no SOS marker")


def _build_decode_table(bits: np.ndarray, vals: np.ndarray):
    """Canonical Huffman code -> {(length, code): symbol} (T.81 C.2)."""
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(int(bits[length])):
            table[(length, code)] = int(vals[k])
            code += 1
            k += 1
        code <<= 1
    return table


def decode_coefficients(h: JpegHeader):
    """Entropy-decode the scan (T.81 F.2.2).  Returns per component an int32 array [blocks_h, blocks_w, 64] of quantised
    coefficients in natural order (MCU-padded block grid) plus (hmax, vmax)."""
    # un-stuff: FF 00 -> FF; restart markers are not produced by cv2.imwrite defaults
    raw = bytearray()
    d = h.data
    i = 0
    while i < len(d):
        b = d[i]
        if b == 0xFF:
            nxt = d[i + 1] if i + 1 < len(d) else 0
            if nxt == 0:
                raw.append(0xFF)
                i += 2
                continue
            if 0xD0 <= nxt <= 0xD7:
                raise ValueError("restart markers are not covered")
            break
        raw.append(b)
        i += 1
    raw += b"\x00" * 8
    hmax = max(c[1] for c in h.components)
    vmax = max(c[2] for c in h.components)
    mcus_w = -(-h.width // (8 * hmax))
    mcus_h = -(-h.height // (8 * vmax))
    coefs = [np.zeros((mcus_h * c[2], mcus_w * c[1], 64), dtype=np.int32) for c in h.components]
    tables = {k: _build_decode_table(*v) for k, v in h.huff.items()}
    pos = 0                                             # bit position

    def get_bits(n):
        nonlocal pos
        v = 0
        for _ in range(n):
            v = (v << 1) | ((raw[pos >> 3] >> (7 - (pos & 7))) & 1)
            pos += 1
        return v

    def decode_symbol(tab):
        nonlocal pos
        code = 0
        for length in range(1, 17):
            code = (code << 1) | ((raw[pos >> 3] >> (7 - (pos & 7))) & 1)
            pos += 1
            s = tab.get((length, code))
            if s is not None:
                return s
        raise ValueError("bad Huffman code")

    def extend(v, t):
        return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1

    pred = [0] * len(h.components)
    for my in range(mcus_h):
        for mx in range(mcus_w):
            for (ci, td, ta) in h.scan:
                _, ch, cv, _ = h.components[ci]
                for by in range(cv):
                    for bx in range(ch):
                        blk = coefs[ci][my * cv + by, mx * ch + bx]
                        t = decode_symbol(tables[(0, td)])
                        pred[ci] += extend(get_bits(t), t)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = decode_symbol(tables[(1, ta)])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break                  # EOB
                                k += 16                    # ZRL
                                continue
                            k += r
                            blk[ZIGZAG[k]] = extend(get_bits(s), s)
                            k += 1
    return coefs, hmax, vmax


# ------------------------------------------------------------------------------------------------ jidctint.c (islow)
CONST_BITS, PASS1_BITS = 13, 2
F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _idct_1d(d, shift):
    """One pass of jpeg_idct_islow over the last axis of int64 array d [..., 8]; DESCALE by `shift`."""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    z2, z3 = d[..., 0], d[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175875602
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298631336, tmp1 * F_2_053119869, tmp2 * F_3_072711026, tmp3 * F_1_501321110
    z1, z2, z3, z4 = z1 * (-F_0_899976223), z2 * (-F_2_562915447), z3 * (-F_1_961570560) + z5, z4 * (-F_0_390180644) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    rnd = 1 << (shift - 1)
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
    return (out + rnd) >> shift


def range_limit_idct(x):
    """sample_range_limit + CENTERJSAMPLE indexed by (x & RANGE_MASK) (jdmaster.c prepare_range_limit_table)."""
    idx = x & 1023
    return np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896))).astype(np.uint8)


def idct_blocks(coefs: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """[bh, bw, 64] quantised coefficients -> uint8 samples [bh*8, bw*8]."""
    bh, bw, _ = coefs.shape
    d = (coefs.astype(np.int64) * qt.astype(np.int64)).reshape(bh, bw, 8, 8)
    ws = _idct_1d(d.transpose(0, 1, 3, 2), CONST_BITS - PASS1_BITS)        # pass 1: columns (last axis = row index)
    out = _idct_1d(ws.transpose(0, 1, 3, 2), CONST_BITS + PASS1_BITS + 3)  # pass 2: rows
    return range_limit_idct(out).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


# ------------------------------------------------------------------------------------------------ jdsample.c
def h2v2_fancy_upsample(c: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """c: chroma samples, real area [rows, cols] (downsampled size) -> [2*rows, 2*cols] uint8."""
    x = c[:rows, :cols].astype(np.int64)
    above = np.vstack([x[:1], x[:-1]])          # row -1 := row 0 (context row duplication, jdmainct.c)
    below = np.vstack([x[1:], x[-1:]])          # row `rows` := last real row
    out = np.zeros((2 * rows, 2 * cols), dtype=np.int64)
    for v, other in ((0, above), (1, below)):
        s = 3 * x + other                        # "colsum" of every chroma column
        if cols == 1:
            out[v::2, 0] = (s[:, 0] * 4 + 8) >> 4
            out[v::2, 1] = (s[:, 0] * 4 + 7) >> 4
            continue
        last = np.hstack([s[:, :1], s[:, :-1]])
        nxt = np.hstack([s[:, 1:], s[:, -1:]])
        even = (3 * s + last + 8) >> 4
        odd = (3 * s + nxt + 7) >> 4
        even[:, 0] = (s[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (s[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------ jdcolor.c
def ycc_to_bgr(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    fix = lambda v: int(v * 65536 + 0.5)
    y = y.astype(np.int64)
    xb, xr = cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((fix(1.40200) * xr + 32768) >> 16)
    b = y + ((fix(1.77200) * xb + 32768) >> 16)
    g = y + (((-fix(0.34414)) * xb + 32768 + (-fix(0.71414)) * xr) >> 16)
    return np.clip(np.stack([b, g, r], axis=-1), 0, 255).astype(np.uint8)


def decode(buf: bytes) -> np.ndarray:
    """Baseline JPEG bytes -> uint8 [H, W, 3] in BGR order, bit-identical to cv2.imdecode(..., IMREAD_COLOR)."""
    h = parse(buf)
    coefs, hmax, vmax = decode_coefficients(h)
    planes = []
    for ci, (cid, ch, cv, tq) in enumerate(h.components):
        samples = idct_blocks(coefs[ci], h.qt[tq])
        if (ch, cv) == (hmax, vmax):
            planes.append(samples[: h.height, : h.width])
        elif (ch, cv) == (1, 1) and (hmax, vmax) == (2, 2):
            rows, cols = -(-h.height // 2), -(-h.width // 2)
            planes.append(h2v2_fancy_upsample(samples, rows, cols)[: h.height, : h.width])
        else:
            raise ValueError(f"sampling {ch}x{cv} of {hmax}x{vmax}: only 4:4:4 and 4:2:0 are covered")
    if len(planes) == 1:
        return np.repeat(planes[0][:, :, None], 3, axis=2)
    return ycc_to_bgr(planes[0], planes[1], planes[2])
