"""Batched baseline-JPEG decode on the GPU: the drop-in for the `cv2.imread` of the face crops
(reference src/get_prob_video.py:95; the files are written by src/data/get_face_images.py:60 with cv2.imwrite defaults).

The host does what is inherently serial and tiny -- walking the ~20 marker segments of every file (T.81 B.2: SOF0, DQT, DHT,
SOS) -- and hands the batch to `avcer_jpeg_decode` (csrc/jpeg.cu): byte un-stuffing as a stream compaction, Huffman
decoding with one thread per image, libjpeg-turbo's integer IDCT, fancy chroma up-sampling and YCbCr -> BGR conversion,
bit-identical to cv2.imread.  Covered: baseline sequential DCT, 8 bit, three components, 4:2:0 or 4:4:4, no restart markers
(everything cv2.imwrite produces by default).  Other files raise `UnsupportedJpeg`; `config.set_jpeg_decoder("cv2")` hands
decoding back to the host (upstream of the accelerated path, as in the reference).
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], dtype=np.int64)

IMAGE_DTYPE = np.dtype([("data_off", "<i8"), ("data_len", "<i8"), ("coef_off", "<i8"), ("plane_off", "<i8"), ("out_off", "<i8"),
                        ("width", "<i4"), ("height", "<i4"), ("mcus_w", "<i4"), ("mcus_h", "<i4"), ("hs", "<i4"), ("qt_y", "<i4"),
                        ("qt_c", "<i4"), ("src_shift", "<i4")], align=True)          # avcer_jpeg_image


PROFILE = None          # measurement aid: a list receiving (event before, event after) around every avcer_jpeg_decode call


class UnsupportedJpeg(ValueError):
    pass


class ParsedJpeg:
    __slots__ = ("width", "height", "hs", "qt_y", "qt_c", "huff_bits", "huff_vals", "data", "scan_start", "dims_off")


def _u16(b: bytes, i: int) -> int:
    return (b[i] << 8) | b[i + 1]


def parse(buf: bytes) -> ParsedJpeg:
    """Marker walk of one file.  Returns frame size, sampling (hs), the two quantisation tables in natural order, the four
    Huffman tables (DC / AC of luma, DC / AC of chroma) and the entropy-coded segment as it sits in the file."""
    if len(buf) < 4 or buf[0] != 0xFF or buf[1] != 0xD8:
        raise UnsupportedJpeg("not a JPEG file (no SOI marker)")
    qt, huff, comps, scan = {}, {}, [], None
    width = height = 0
    i = 2
    n = len(buf)
    while i + 4 <= n:
        if buf[i] != 0xFF:
            raise UnsupportedJpeg(f"marker expected at byte {i}")
        m = buf[i + 1]
        if m == 0xFF:
            i += 1
            continue
        seg_len = _u16(buf, i + 2)
        seg = buf[i + 4: i + 2 + seg_len]
        if m == 0xC0:
            if seg[0] != 8:
                raise UnsupportedJpeg("only 8-bit samples")
            height, width = _u16(seg, 1), _u16(seg, 3)
            dims_off = i + 5                                       # file offset of the height / width bytes (2 + 2)
            comps = [(seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c]) for c in range(seg[5])]
        elif 0xC1 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise UnsupportedJpeg(f"SOF marker 0x{m:02x}: only baseline sequential DCT (SOF0) is decoded on the GPU")
        elif m == 0xDB:
            j = 0
            while j < len(seg):
                if seg[j] >> 4:
                    raise UnsupportedJpeg("16-bit quantisation tables")
                t = np.zeros(64, dtype=np.uint16)
                t[ZIGZAG] = np.frombuffer(seg, dtype=np.uint8, count=64, offset=j + 1)
                qt[seg[j] & 15] = t
                j += 65
        elif m == 0xC4:
            j = 0
            while j < len(seg):
                bits = np.frombuffer(seg, dtype=np.uint8, count=16, offset=j + 1)
                cnt = int(bits.sum())
                vals = np.zeros(256, dtype=np.uint8)
                vals[:cnt] = np.frombuffer(seg, dtype=np.uint8, count=cnt, offset=j + 17)
                huff[(seg[j] >> 4, seg[j] & 15)] = (bits.copy(), vals)
                j += 17 + cnt
        elif m == 0xDD:
            if _u16(seg, 0) != 0:
                raise UnsupportedJpeg("restart intervals")
        elif m == 0xDA:
            ids = [c[0] for c in comps]
            scan = [(ids.index(seg[1 + 2 * k]), seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(seg[0])]
            i += 2 + seg_len
            break
        i += 2 + seg_len
    if scan is None or not comps:
        raise UnsupportedJpeg("no SOF0 / SOS marker")
    if len(comps) != 3 or [s[0] for s in scan] != [0, 1, 2]:
        raise UnsupportedJpeg(f"{len(comps)} components: only three-component YCbCr files are decoded on the GPU")
    (_, yh, yv, yq), (_, bh, bv, bq), (_, rh, rv, rq) = comps
    if (bh, bv, rh, rv) != (1, 1, 1, 1) or yh != yv or yh not in (1, 2) or bq != rq or scan[1][1:] != scan[2][1:]:
        raise UnsupportedJpeg(f"sampling {yh}x{yv} / {bh}x{bv} / {rh}x{rv}: only 4:2:0 and 4:4:4 are decoded on the GPU")
    end = buf.rfind(b"\xff\xd9")
    out = ParsedJpeg()
    out.width, out.height, out.hs = width, height, yh
    out.qt_y, out.qt_c = qt[yq], qt[bq]
    tabs = [huff[(0, scan[0][1])], huff[(1, scan[0][2])], huff[(0, scan[1][1])], huff[(1, scan[1][2])]]
    out.huff_bits = np.stack([t[0] for t in tabs])
    out.huff_vals = np.stack([t[1] for t in tabs])
    out.data = buf[i: end if end >= i else n]            # entropy-coded segment, still byte-stuffed (the device removes the stuffing)
    out.scan_start = i
    out.dims_off = dims_off
    return out


def unstuff(data: bytes) -> bytes:
    """Host restatement of the device's un-stuffing pass (tests): FF 00 -> FF."""
    return bytes(data).replace(b"\xff\x00", b"\xff")


_header_cache = {}       # header bytes (SOI .. end of the SOS segment) -> ParsedJpeg without data
_stage_bufs = {}         # device -> pinned uint8 staging buffer (grow-only)
_upload_events = {}      # data_ptr of a pinned host buffer -> event recorded after the last H2D copy out of it


def wait_uploaded(buf: torch.Tensor) -> None:
    """Block until the last decode_packed() that read `buf` has copied it to the device (the buffer may be refilled)."""
    ev = _upload_events.get(buf.data_ptr())
    if ev is not None:
        ev.synchronize()


def _staging(dev: torch.device, nbytes: int) -> torch.Tensor:
    buf = _stage_bufs.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = _stage_bufs[dev] = torch.empty(max(nbytes, 1 << 20) * 5 // 4, dtype=torch.uint8, pin_memory=True)
    else:
        wait_uploaded(buf)                        # the previous batch's copy has left the buffer
    return buf


def _parse_cached(f, last):
    """Files written by one encoder with one setting share their whole header byte for byte except the four height / width
    bytes of SOF0 (cv2.imwrite: 623 bytes; the face crops of a clip all differ in size): the marker walk runs once per
    distinct header family, every other file costs two slice comparisons.  Returns ((header bytes, ParsedJpeg of the family's
    first file, Huffman table key), height, width).  `f`: bytes or a memoryview."""
    if last is not None:
        hdr, p = last[0][0], last[0][1]
        d, n = p.dims_off, len(hdr)
        if f[:d] == hdr[:d] and f[d + 4:n] == hdr[d + 4:n]:
            return last[0], (f[d] << 8) | f[d + 1], (f[d + 2] << 8) | f[d + 3]
    p = parse(bytes(f))
    hdr = bytes(f[:p.scan_start])
    key = hdr[:p.dims_off] + hdr[p.dims_off + 4:]
    hit = _header_cache.get(key)
    if hit is None:
        if len(_header_cache) > 256:
            _header_cache.clear()
        p.data = None
        hit = _header_cache[key] = (hdr, p, p.huff_bits.tobytes() + p.huff_vals.tobytes())
    return hit, p.height, p.width


class PendingStatus:
    """Device-side decode status of one decode_batch(..., defer_status=True) call: check() synchronises and raises
    UnsupportedJpeg like the immediate form would have."""

    def __init__(self, statuses, base: int = 0):
        self.statuses, self.base = statuses, base

    def check(self) -> None:
        for status, idx in self.statuses:
            s = int(status.item())
            if s > 0:
                raise UnsupportedJpeg(f"corrupt entropy-coded data in image {self.base + int(idx[s - 1])} of the batch")
            if s < 0:
                raise UnsupportedJpeg(f"restart markers in image {self.base + int(idx[-s - 1])} of the batch")


def decode_batch(files: Sequence[bytes], device, align_out: int = 16, defer_status: bool = False):
    """Decode a batch of JPEG byte strings.  Returns (flat uint8 device buffer, byte offsets [n], heights [n], widths [n]):
    image i is out[offsets[i] : offsets[i] + h*w*3] viewed as [h, w, 3] in BGR order -- what cv2.imread returns, and the
    ragged layout Engine.vs_forward_ragged / avcer_preprocess_u8 consume.  `align_out`: alignment of every image's offset.
    defer_status: do not synchronise on the decoder's status word; a fifth return value (PendingStatus) checks it later, so
    the host can prepare the next batch while this one decodes.
    The files are packed into one pinned staging buffer (one host copy) and handed to decode_packed."""
    dev = torch.device(device)
    n = len(files)
    sizes = np.fromiter((len(f) for f in files), dtype=np.int64, count=n)
    padded = (sizes + 15) // 16 * 16
    file_off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
    total = int(padded.sum())
    stage = _staging(dev, total + 16)
    stage_np = stage.numpy()
    for f, o, z in zip(files, file_off.tolist(), sizes.tolist()):
        stage_np[o:o + z] = np.frombuffer(f, dtype=np.uint8, count=z)
    return decode_packed(stage, file_off, sizes, dev, align_out, defer_status)


def decode_packed(stage: torch.Tensor, file_off: np.ndarray, file_size: np.ndarray, device, align_out: int = 16,
                  defer_status: bool = False):
    """decode_batch for files that already sit in ONE pinned host buffer (`stage`; file i = stage[file_off[i] : + file_size[i]],
    e.g. filled by avcer_read_files): the whole buffer goes to the device in one copy and every image's entropy-coded segment
    is addressed in place (avcer_jpeg_image.data_off = its position rounded down to a word, src_shift = the remainder) -- no
    per-file staging copy.  wait_uploaded(stage) tells when the buffer may be refilled."""
    dev = torch.device(device)
    n = len(file_off)
    if n == 0:
        empty = (torch.zeros(16, dtype=torch.uint8, device=dev), np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32))
        return empty + (PendingStatus([]),) if defer_status else empty
    mv = memoryview(stage.numpy())
    heads, starts, ends, hh, ww = [], [], [], [], []
    last = None
    for o, z in zip(file_off.tolist(), file_size.tolist()):
        f = mv[o:o + z]
        last = _parse_cached(f, last)
        head = last[0]
        heads.append(head)
        hh.append(last[1])
        ww.append(last[2])
        end = z - 2 if f[-2:] == b"\xff\xd9" else bytes(f).rfind(b"\xff\xd9")      # EOI: normally the last two bytes
        starts.append(len(head[0]))
        ends.append(end if end >= len(head[0]) else z)
    heights = np.array(hh, dtype=np.int32)
    widths = np.array(ww, dtype=np.int32)
    hs = np.array([h[1].hs for h in heads], dtype=np.int32)
    starts, ends = np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64)
    sizes = heights.astype(np.int64) * widths * 3
    padded = (sizes + align_out - 1) // align_out * align_out
    offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    out = torch.empty(int(padded.sum()) + 16, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    cur = torch.cuda.current_stream(dev)
    stream = cur.cuda_stream
    total = int(file_off[-1] + (file_size[-1] + 15) // 16 * 16) + 8
    total = min(total, stage.numel())
    raw = torch.empty(total, dtype=torch.uint8, device=dev)
    raw.copy_(stage[:total], non_blocking=True)
    ev = _upload_events.get(stage.data_ptr())
    if ev is None:
        ev = _upload_events[stage.data_ptr()] = torch.cuda.Event()
    ev.record(cur)
    data = torch.empty_like(raw)
    src = np.asarray(file_off, dtype=np.int64) + starts               # first byte of every entropy-coded segment in `raw`
    # images that share their Huffman tables go in one launch (cv2.imwrite always uses the Annex-K tables)
    groups = {}
    for i, h in enumerate(heads):
        groups.setdefault(h[2], []).append(i)
    statuses = []
    for idx in groups.values():
        idx = np.asarray(idx)
        m = len(idx)
        qtabs, qindex = [], {}

        def qslot(t):
            key = t.tobytes()
            if key not in qindex:
                qindex[key] = len(qtabs)
                qtabs.append(t)
            return qindex[key]

        qy = np.array([qslot(heads[i][1].qt_y) for i in idx], dtype=np.int32)
        qc = np.array([qslot(heads[i][1].qt_c) for i in idx], dtype=np.int32)
        w, h, s = widths[idx].astype(np.int64), heights[idx].astype(np.int64), hs[idx].astype(np.int64)
        mw, mh = -(-w // (8 * s)), -(-h // (8 * s))
        blocks = mw * mh * (s * s + 2)
        imgs = np.zeros(m, dtype=IMAGE_DTYPE)
        imgs["data_off"] = src[idx] & ~np.int64(3)
        imgs["src_shift"] = (src[idx] & 3).astype(np.int32)           # src_shift
        imgs["data_len"] = ends[idx] - starts[idx]
        imgs["coef_off"] = np.concatenate([[0], np.cumsum(blocks)[:-1]])
        imgs["plane_off"] = imgs["coef_off"] * 64
        imgs["out_off"] = offsets[idx]
        imgs["width"], imgs["height"], imgs["mcus_w"], imgs["mcus_h"], imgs["hs"] = w, h, mw, mh, s
        imgs["qt_y"], imgs["qt_c"] = qy, qc
        prefix = np.concatenate([[0], np.cumsum(w * h)[:-1]]).astype(np.int64)
        coef_total, pix = int(blocks.sum()), int((w * h).sum())
        dlens = torch.empty(m, dtype=torch.int64, device=dev)
        meta = torch.from_numpy(imgs.view(np.uint8).reshape(-1)).to(dev)
        p0 = heads[idx[0]][1]
        bits = torch.from_numpy(p0.huff_bits.reshape(-1).copy()).to(dev)
        vals = torch.from_numpy(p0.huff_vals.reshape(-1).copy()).to(dev)
        qt = torch.from_numpy(np.stack(qtabs).astype(np.uint16).view(np.int16)).to(dev)
        pre = torch.from_numpy(prefix).to(dev)
        coefs = torch.empty(coef_total * 64, dtype=torch.int16, device=dev)
        planes = torch.empty(coef_total * 64 + 8, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if PROFILE is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = lib.avcer_jpeg_decode(raw.data_ptr(), meta.data_ptr(), m, bits.data_ptr(), vals.data_ptr(), qt.data_ptr(),
                                   pre.data_ptr(), coef_total, pix, data.data_ptr(), dlens.data_ptr(), coefs.data_ptr(), planes.data_ptr(),
                                   out.data_ptr(), status.data_ptr(), ctypes.c_void_p(stream))
        _lib.check(rc)
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((e0, e1))
        statuses.append((status, idx))
    pending = PendingStatus(statuses)
    if defer_status:
        return out, offsets, heights, widths, pending
    pending.check()
    return out, offsets, heights, widths


def decode_images(files: Sequence[bytes], device) -> List[torch.Tensor]:
    """Convenience: one uint8 [h, w, 3] BGR device tensor per file (views into one buffer)."""
    out, off, hs, ws = decode_batch(files, device)
    return [out[int(o): int(o) + int(h) * int(w) * 3].view(int(h), int(w), 3) for o, h, w in zip(off, hs, ws)]
