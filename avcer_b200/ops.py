"""Tensor-level wrappers over the C ABI (torch tensors in, torch tensors out).

PyTorch is plumbing here: it owns device memory and streams; every computation below is a
hand-written kernel of libavcer_b200.so.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, F32, ContractDesc
from ._lib import check as _check

__all__ = [
    "F32", "BF16", "ACT_NONE", "ACT_RELU", "ACT_GELU", "dtype_code", "torch_dtype", "contract", "conv2d_nhwc",
    "linear", "preprocess", "fuse_compound", "weight_search_confusion", "softmax7", "window_to_frame_mean", "gather_rows", "maxpool3x3s2",
    "stem_pool", "stem_pool_u8", "subsample_rows", "avgpool", "small_linear", "lstm_cell", "gru_cell", "split_bf16x3", "sinc_resample_bank", "pcm16_to_mono", "audio_normalize_windows", "w2v_conv0_ln_gelu", "w2v_conv0_tc", "det_stem", "det_stem_tc", "maxpool3x3s2p1", "upsample_add", "det_decode", "nearest_source_index", "layernorm", "add_rows",
    "attention", "maxpool1d5_relu", "avgpool1d_relu", "cast", "sm_limit",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Launch accounting (bench.py's `gpu_launches`) and optional per-kernel CUDA-event timing on the
# launching stream (bench.py's roofline): PROFILE is None or a list receiving (tag, work, ev0, ev1).
STATS = {"launches": 0}
PROFILE = None


def check(rc: int) -> None:
    """Every C-ABI call below launches exactly one kernel (the _Timed ones count themselves)."""
    STATS["launches"] += 1
    _check(rc)


CONTRACT_KERNELS = {1: "simt_contract_kernel", 2: "tc_gemm_kernel<64>", 3: "tc_gemm_kernel<128>", 4: "tc_gemm_kernel<256>",
                    5: "tc_gemm_kernel<f32 out>", 6: "tc_gemm2_kernel<OUT_TMA,256,ring>", 7: "tc_gemm2_kernel<OUT_TMA_RES,256,ring>",
                    8: "tc_gemm2_kernel<OUT_TMA,256,FLAT>", 9: "tc_gemm2_kernel<OUT_TMA_RES,256,FLAT>", 10: "tc_gemm2_kernel<128>",
                    11: "conv3x3_kernel<64>", 12: "conv3x3_kernel<128>", 13: "tc_gemm_kernel<strip>"}


class _Timed:
    __slots__ = ("tag", "work", "e0", "bytes")

    def __init__(self, tag: str, work: float, launches: int = 1):
        STATS["launches"] += launches
        self.tag, self.work, self.e0, self.bytes = tag, work, None, 0.0

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.tag, self.work, self.e0, e1, self.bytes))
        return False


class sm_limit:
    """Context: persistent kernels launched (or captured) inside use `n_sms` SMs instead of the whole device
    (avcer_set_sm_limit); 0 = no limit."""

    def __init__(self, n_sms: int):
        self.n = int(n_sms)

    def __enter__(self):
        if self.n:
            _lib.load()
            for lib in _lib.loaded():                 # every build of the library keeps its own (thread-local) limit
                _check(lib.avcer_set_sm_limit(self.n))
        return self

    def __exit__(self, *exc):
        if self.n:
            for lib in _lib.loaded():
                _check(lib.avcer_set_sm_limit(0))
        return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _L(*tensors: Optional[torch.Tensor]):
    """The library build that matches the operands: libavcer_b200_fp16.so as soon as one of them is float16 (precision
    "fp16"), else libavcer_b200.so (bfloat16 storage; also every fp32 / fp64 / integer kernel)."""
    for t in tensors:
        if t is not None and t.dtype == torch.float16:
            return _lib.load("fp16")
    return _lib.load("bf16")


def dtype_code(dt: torch.dtype) -> int:
    if dt in (torch.bfloat16, torch.float16):          # "the 16-bit storage type of the library" (_L picks the matching build)
        return BF16
    if dt == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {dt}")


def torch_dtype(code: int) -> torch.dtype:
    return torch.bfloat16 if code == BF16 else torch.float32


def _cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.AvcerError(f"{name} must be a CUDA tensor (avcer_b200 has no CPU path)")


# ----------------------------------------------------------------------------------------- contraction
def contract(*, a: torch.Tensor, a_dim: Sequence[int], a_stride: Sequence[int], wt: torch.Tensor,
             bias: Optional[torch.Tensor], out: torch.Tensor, out_stride: Sequence[int], W: int, H: int, NB: int,
             cin: int, cout: int, taps_w: int = 1, taps_h: int = 1, off_w: int = 0, off_h: int = 0,
             tap_h_in_dim4: bool = False, group_cin_shift: int = 0, residual: Optional[torch.Tensor] = None,
             res_stride: Optional[Sequence[int]] = None, act: int = ACT_NONE, res_after_act: bool = False,
             a_offset: int = 0, algo_k: Optional[int] = None, a_strip: bool = False,
             wt_packed: Optional[torch.Tensor] = None, a_step: int = 1, reverse: bool = False) -> torch.Tensor:
    """Generic implicit GEMM (see avcer_contract in include/avcer_b200.h)."""
    _cuda(a, "a")
    d = ContractDesc()
    code = dtype_code(a.dtype)
    d.a = a.data_ptr() + a_offset * a.element_size()
    for i in range(5):
        d.a_dim[i] = int(a_dim[i])
        d.a_stride[i] = int(a_stride[i])
    assert wt.dtype == a.dtype and wt.is_contiguous()
    d.wt = wt.data_ptr()
    d.bias = _ptr(bias)
    d.residual = _ptr(residual)
    d.out = out.data_ptr()
    rs = res_stride if res_stride is not None else out_stride
    for i in range(3):
        d.out_stride[i] = int(out_stride[i])
        d.res_stride[i] = int(rs[i])
    d.W, d.H, d.NB, d.cin, d.cout = W, H, NB, cin, cout
    d.taps_w, d.taps_h, d.off_w, d.off_h = taps_w, taps_h, off_w, off_h
    d.tap_h_in_dim4 = int(tap_h_in_dim4)
    d.group_cin_shift = group_cin_shift
    d.a_strip = int(a_strip)
    d.a_step = int(a_step)
    d.reverse_tiles = int(reverse)
    d.wt_packed = _ptr(wt_packed)
    d.act = act
    d.res_after_act = int(res_after_act)
    d.dtype = code
    d.out_f32 = int(code == BF16 and out.dtype == torch.float32)
    assert wt.dtype == a.dtype and (residual is None or residual.dtype == a.dtype)
    k_real = algo_k if algo_k is not None else taps_w * taps_h * cin
    rows = W * H * NB
    with _Timed("contract_bf16" if code == BF16 else "contract_f32", 2.0 * rows * cout * k_real) as tm:
        _check(_L(a, out).avcer_contract(ctypes.byref(d), _stream()))
        if PROFILE is not None:
            # exact kernel + algorithmic bytes of this launch (operands + output once, at their storage width)
            esz = a.element_size()
            tm.tag = "contract:" + CONTRACT_KERNELS.get(_L(a, out).avcer_last_contract_kernel(), "?")
            tm.bytes = float(esz * (rows * (k_real if taps_w * taps_h == 1 else cin) + cout * taps_w * taps_h * cin + (rows * cout if residual is not None else 0))
                             + out.element_size() * rows * cout)
    return out


def conv2d_nhwc(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], *, kh: int, kw: int, stride: int = 1,
                pad_h: int = 0, pad_w: int = 0, residual: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                out: Optional[torch.Tensor] = None, residual_stride: int = 1, reverse: bool = False) -> torch.Tensor:
    """x: [N,H,W,C] contiguous; wt: [Cout, kh*kw*C] (tap-major, channel-minor).
    residual_stride = s: `residual` is a contiguous [N, Hr, Wr, Cout] tensor read at every s-th pixel (a strided 1x1 conv
    whose residual lives at the input resolution: only the output pixels the next stage samples are computed)."""
    n, h, w, c = x.shape
    cout = wt.shape[0]
    assert wt.shape[1] == kh * kw * c
    ho = (h + 2 * pad_h - kh) // stride + 1
    wo = (w + 2 * pad_w - kw) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), device=x.device, dtype=x.dtype)
    if (kh == 1 and kw == 1 and stride == 1 and residual_stride == 1 and pad_h == 0 and pad_w == 0 and x.is_contiguous() and out.is_contiguous()
            and (residual is None or residual.is_contiguous()) and n * h * w < (1 << 31)):
        # A pointwise conv does not see the image structure: run it as one [n*h*w, C] GEMM so that every M tile is a
        # full 128 rows (a 14x14 image only offers 126 + 70 row boxes, a 7x7 pair 98: 77 % of the MMA rows).
        linear(x.view(n * h * w, c), wt, bias, residual=None if residual is None else residual.view(n * h * w, cout),
               act=act, out=out.view(n * h * w, cout), reverse=reverse)
        return out
    a_step = 1
    if stride == 1:
        a_dim = (c, w, h, n, 1)
        a_stride = (1, c, w * c, h * w * c, n * h * w * c)
    elif kh == 1 and kw == 1 and pad_h == 0 and pad_w == 0:
        a_dim = (c, wo, ho, n, 1)                        # a strided 1x1 conv is a view of every stride-th pixel
        a_stride = (1, stride * c, stride * w * c, h * w * c, n * h * w * c)
    else:
        assert x.dtype != torch.float32, "strided k x k convs: tensor-core path only (TMA traversal stride)"
        a_dim = (c, w, h, n, 1)                          # full-resolution input walked with traversal stride `stride`
        a_stride = (1, c, w * c, h * w * c, n * h * w * c)
        a_step = stride
    op = out.stride(2)                                   # pixel pitch: `out` may be a channel slice of a wider buffer
    assert out.stride(3) == 1 and out.stride(1) == wo * op and out.stride(0) == ho * wo * op
    res_stride = None
    if residual is not None:
        assert residual.is_contiguous() and residual.shape[3] == cout and residual.shape[0] == n
        hr, wr = residual.shape[1], residual.shape[2]
        assert (hr - 1) // residual_stride + 1 == ho and (wr - 1) // residual_stride + 1 == wo
        res_stride = (residual_stride * cout, residual_stride * wr * cout, hr * wr * cout)
    return contract(a=x, a_dim=a_dim, a_stride=a_stride, wt=wt, bias=bias, out=out,
                    out_stride=(op, wo * op, ho * wo * op), W=wo, H=ho, NB=n, cin=c, cout=cout, taps_w=kw,
                    taps_h=kh, off_w=-pad_w, off_h=-pad_h, residual=residual, res_stride=res_stride, act=act, a_step=a_step, reverse=reverse)


def linear(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], *, residual: Optional[torch.Tensor] = None,
           act: int = ACT_NONE, res_after_act: bool = False, out: Optional[torch.Tensor] = None,
           out_dtype: Optional[torch.dtype] = None, reverse: bool = False) -> torch.Tensor:
    """x: [M, K] (row pitch = x.stride(0)), wt: [N, K]; returns [M, N]."""
    m, k = x.shape
    n = wt.shape[0]
    assert x.stride(1) == 1
    ld = x.stride(0)
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=out_dtype or x.dtype)
    big = (max(ld * m, 8) + 7) // 8 * 8
    return contract(a=x, a_dim=(k, m, 1, 1, 1), a_stride=(1, ld, big, big, big), wt=wt, bias=bias, out=out,
                    out_stride=(out.stride(0), 0, 0), W=m, H=1, NB=1, cin=k, cout=n, residual=residual,
                    res_stride=None if residual is None else (residual.stride(0), 0, 0), act=act,
                    res_after_act=res_after_act, reverse=reverse)


# ----------------------------------------------------------------------------------------- K1
PAD_H, PAD_W = 232, 240      # zero-bordered stem input: rows x row pitch (pixels); 240 px = 15 x 128 B


def preprocess(src: torch.Tensor, n: int, dst: torch.Tensor, layout: int, *, offsets: Optional[torch.Tensor] = None,
               heights: Optional[torch.Tensor] = None, widths: Optional[torch.Tensor] = None,
               maps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src: flat uint8 device buffer of BGR HWC crops. layout 0: f32 NCHW, 1: bf16 NHWC4 padded, 2: f32 NHWC4 padded."""
    _cuda(src, "src")
    lib = _L(dst)
    if heights is not None and maps is None:
        maps = torch.empty((n, 2, 224), device=src.device, dtype=torch.int16)
        check(lib.avcer_preprocess_maps(heights.data_ptr(), widths.data_ptr(), n, maps.data_ptr(), _stream()))
    # algorithmic bytes per frame: u8 in + bf16 NHWC C=3 out (fp32: 4 B/elem), SURVEY.md section 8d
    work = n * (150528 + 150528 * (2 if layout == 1 else 4))
    with _Timed("preprocess", work):
        _check(lib.avcer_preprocess_u8(src.data_ptr(), _ptr(offsets), _ptr(heights), _ptr(widths), _ptr(maps), n,
                                      dst.data_ptr(), layout, _stream()))
    return dst


# ----------------------------------------------------------------------------------------- K4 + glue
def fuse_compound(p_vs: torch.Tensor, p_vd: torch.Tensor, p_a: torch.Tensor, weights_1, weights_2,
                  ce_weights_type: bool, ce_mask: bool, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns int64 labels [4, n] = (AV, VS, VD, A) compound-expression argmax.
    `labels` may be a column slice [4, n] of a wider [4, total] buffer (row pitch labels.stride(0))."""
    _cuda(p_vs, "p_vs")
    n = p_vs.shape[0]
    assert p_vs.shape == p_vd.shape == p_a.shape == (n, 7)
    assert p_vs.dtype == p_vd.dtype == p_a.dtype
    for t in (p_vs, p_vd, p_a):
        assert t.is_contiguous()
    if labels is None:
        labels = torch.empty((4, n), device=p_vs.device, dtype=torch.int64)
    assert labels.dtype == torch.int64 and tuple(labels.shape) == (4, n) and (n <= 1 or labels.stride(1) == 1)
    pitch = labels.stride(0) if n > 0 else 0
    w1 = w2 = None
    if weights_1:
        w1a = np.ascontiguousarray(np.asarray(weights_1, dtype=np.float64).reshape(3, 7))
        w2a = np.ascontiguousarray(np.asarray(weights_2, dtype=np.float64).reshape(3))
        w1 = w1a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        w2 = w2a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib = _lib.load()
    fn = lib.avcer_fuse_compound if p_vs.dtype == torch.float32 else lib.avcer_fuse_compound_f64
    with _Timed("fuse_compound", n * 116.0):      # 84 B in + 32 B of int64 labels out per frame
        _check(fn(p_vs.data_ptr(), p_vd.data_ptr(), p_a.data_ptr(), n, w1, w2, int(bool(ce_weights_type)),
                 int(bool(ce_mask)), labels.data_ptr(), pitch, _stream()))
    return labels


def weight_search_confusion(preds: torch.Tensor, gt: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """preds [M, n, 7] f64, gt [n] int32, weights [W, M, 7] f64 (all CUDA) -> confusion counts [W, 7, 7] int64
    (rows = ground truth, columns = fused argmax) for every candidate weight set."""
    _cuda(preds, "preds")
    m, n, c = preds.shape
    w = weights.shape[0]
    assert c == 7 and weights.shape == (w, m, 7) and gt.shape == (n,)
    assert preds.dtype == torch.float64 and weights.dtype == torch.float64 and gt.dtype == torch.int32
    assert preds.is_contiguous() and weights.is_contiguous() and gt.is_contiguous()
    cm = torch.zeros((w, 7, 7), device=preds.device, dtype=torch.int64)
    check(_lib.load().avcer_weight_search_confusion(preds.data_ptr(), m, n, gt.data_ptr(), weights.data_ptr(), w,
                                                    cm.data_ptr(), _stream()))
    return cm


def softmax7(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row softmax over the first 7 columns of x [n, >=7] (f32 or f64)."""
    _cuda(x, "x")
    n, ld = x.shape
    assert x.is_contiguous()
    y = out if out is not None else torch.empty((n, 7), device=x.device, dtype=x.dtype)
    assert tuple(y.shape) == (n, 7) and y.is_contiguous() and y.dtype == x.dtype
    lib = _lib.load()
    fn = lib.avcer_softmax7 if x.dtype == torch.float32 else lib.avcer_softmax7_f64
    check(fn(x.data_ptr(), n, ld, y.data_ptr(), _stream()))
    return y


def window_to_frame_mean(logits: torch.Tensor, f_lo: torch.Tensor, f_hi: torch.Tensor, n_frames: int) -> torch.Tensor:
    """pandas groupby.mean in the dtype of `logits` (float32 or float64), see avcer_window_to_frame_mean."""
    _cuda(logits, "logits")
    n_win, ncls = logits.shape
    assert logits.dtype in (torch.float32, torch.float64) and logits.is_contiguous()
    out = torch.empty((n_frames, ncls), device=logits.device, dtype=logits.dtype)
    fn = _lib.load().avcer_window_to_frame_mean if logits.dtype == torch.float32 else _lib.load().avcer_window_to_frame_mean_f64
    check(fn(logits.data_ptr(), n_win, ncls, f_lo.data_ptr(), f_hi.data_ptr(), n_frames, out.data_ptr(), _stream()))
    return out


def gather_rows(src: torch.Tensor, index: Optional[torch.Tensor], n_out: int, perm: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = src[index[i]] (a zero row for index -1; identity when index is None), columns optionally permuted.
    `out`: optional contiguous [n_out, ncols] destination (e.g. a slice of an all-gather send buffer)."""
    _cuda(src, "src")
    ncols = src.shape[1] if perm is None else perm.numel()
    assert perm is None or ncols == src.shape[1]
    assert src.dtype in (torch.float32, torch.float64) and src.is_contiguous()
    if out is None:
        out = torch.empty((n_out, ncols), device=src.device, dtype=src.dtype)
    assert tuple(out.shape) == (n_out, ncols) and out.is_contiguous() and out.dtype == src.dtype
    fn = _lib.load().avcer_gather_rows if src.dtype == torch.float32 else _lib.load().avcer_gather_rows_f64
    check(fn(src.data_ptr(), _ptr(index), n_out, ncols, _ptr(perm), out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------------------- small layers
def stem_pool(x: torch.Tensor, wt_packed: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused ResNet stem + max-pool: x [n,232,240,4] bf16 zero-bordered crops -> [n,55,55,64] bf16.
    `out`: optional channel slice [n,55,55,64] of a wider contiguous [n,55,55,C] buffer (pixel pitch C)."""
    _cuda(x, "x")
    n = x.shape[0]
    assert x.dtype in (torch.bfloat16, torch.float16) and x.is_contiguous() and tuple(x.shape[1:]) == (PAD_H, PAD_W, 4)
    if out is None:
        out = torch.empty((n, 55, 55, 64), device=x.device, dtype=x.dtype)
    pitch = out.stride(2)
    assert out.dtype == x.dtype and tuple(out.shape) == (n, 55, 55, 64) and out.stride(3) == 1
    assert out.stride(1) == 55 * pitch and out.stride(0) == 55 * 55 * pitch
    with _Timed("contract_bf16", 2.0 * n * 112 * 112 * 64 * 147):        # 7x7x3 real taps
        _check(_L(x).avcer_stem_pool_ld(x.data_ptr(), wt_packed.data_ptr(), bias.data_ptr(), n, out.data_ptr(), pitch, _stream()))
    return out


def stem_pool_u8(crops: torch.Tensor, wt_packed: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1 + stem + max-pool in one kernel: crops uint8 [n,224,224,3] (BGR, packed 224x224) -> [n,55,55,64] bf16, bit-identical
    to preprocess(layout 1) + stem_pool.  `out` as in stem_pool."""
    _cuda(crops, "crops")
    n = crops.shape[0]
    assert crops.dtype == torch.uint8 and crops.is_contiguous() and tuple(crops.shape[1:]) == (224, 224, 3)
    if out is None:
        out = torch.empty((n, 55, 55, 64), device=crops.device, dtype=torch.bfloat16)
    pitch = out.stride(2)
    assert out.dtype in (torch.bfloat16, torch.float16) and tuple(out.shape) == (n, 55, 55, 64) and out.stride(3) == 1
    assert out.stride(1) == 55 * pitch and out.stride(0) == 55 * 55 * pitch
    # K1's algorithmic bytes are part of this launch; the tensor work is the stem's (7x7x3 real taps)
    with _Timed("contract_bf16", 2.0 * n * 112 * 112 * 64 * 147):
        _check(_L(out).avcer_stem_pool_u8(crops.data_ptr(), wt_packed.data_ptr(), bias.data_ptr(), n, out.data_ptr(), pitch, _stream()))
    return out


def subsample_rows(x: torch.Tensor, stride: int, out: torch.Tensor) -> torch.Tensor:
    """x: [n,h,w,c] contiguous; out: [n*ho*wo, c] view (row pitch out.stride(0) >= c) receiving x[:, ::stride, ::stride]."""
    _cuda(x, "x")
    n, h, w, c = x.shape
    assert x.is_contiguous() and out.stride(1) == 1 and out.shape[1] == c and out.dtype == x.dtype
    assert out.shape[0] == n * ((h - 1) // stride + 1) * ((w - 1) // stride + 1)
    check(_L(x).avcer_subsample_rows(x.data_ptr(), n, h, w, c, stride, out.data_ptr(), out.stride(0), dtype_code(x.dtype), _stream()))
    return out


def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty((n, (h - 3) // 2 + 1, (w - 3) // 2 + 1, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_maxpool3x3s2(x.data_ptr(), n, h, w, c, y.data_ptr(), dtype_code(x.dtype), _stream()))
    return y


# ----------------------------------------------------------------------------------------- face detector layers
def det_stem(frames: torch.Tensor, wt: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype, rgb: bool = False) -> torch.Tensor:
    """frames: uint8 [n,h,w,3] video frames (B,G,R; rgb: R,G,B) -> relu(conv7x7/2(frames - mean) + bias) [n,ceil(h/2),ceil(w/2),64]."""
    _cuda(frames, "frames")
    n, h, w, c = frames.shape
    assert frames.dtype == torch.uint8 and c == 3 and frames.is_contiguous() and wt.shape == (147, 64) and wt.dtype == torch.float32
    y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 64), device=frames.device, dtype=dtype)
    with _Timed("det_stem", 2.0 * y.numel() * 147):
        _check(_L(y).avcer_det_stem(frames.data_ptr(), n, h, w, int(rgb), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), dtype_code(dtype), _stream()))
    return y


def det_stem_tc(frames: torch.Tensor, wt: torch.Tensor, wt_packed: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype,
                rgb: bool = False) -> torch.Tensor:
    """The same layer on the tensor cores (16-bit storage): avcer_det_prepare writes the zero-bordered NHWC4 copy of the
    frames (mean-subtracted pixels are exact integers), then one strip-mode contraction per band of 128 output columns
    (wt [64, 7*32] / wt_packed: the VS stem's layouts, weights.pack_retinaface)."""
    _cuda(frames, "frames")
    n, h, w, c = frames.shape
    assert frames.dtype == torch.uint8 and c == 3 and frames.is_contiguous() and dtype in (torch.bfloat16, torch.float16)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    bands = (wo + 127) // 128
    hp = h + 6
    wp = (max(w + 6, 256 * (bands - 1) + 272) + 15) // 16 * 16          # a band's strip is 17 x 64 elements = 272 pixels
    x = torch.empty((n, hp, wp, 4), device=frames.device, dtype=dtype)
    check(_L(x).avcer_det_prepare(frames.data_ptr(), n, h, w, int(rgb), hp, wp, x.data_ptr(), _stream()))
    y = torch.empty((n, ho, wo, 64), device=frames.device, dtype=dtype)
    for s in range(bands):
        bw = min(128, wo - 128 * s)
        contract(a=x, a_offset=256 * s * 4, a_dim=(64, 17, ho, n, 7), a_stride=(1, 64, 2 * wp * 4, hp * wp * 4, wp * 4), wt=wt, bias=bias,
                 out=y[:, :, 128 * s:, :], out_stride=(64, wo * 64, ho * wo * 64), W=bw, H=ho, NB=n, cin=32, cout=64, taps_w=1, taps_h=7,
                 tap_h_in_dim4=True, act=ACT_RELU, algo_k=147, a_strip=True, wt_packed=wt_packed)
    return y


def maxpool3x3s2p1(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    assert x.is_contiguous()
    y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_maxpool3x3s2p1(x.data_ptr(), n, h, w, c, y.data_ptr(), dtype_code(x.dtype), _stream()))
    return y


def nearest_source_index(n_in: int, n_out: int) -> torch.Tensor:
    """Source index of every output position of F.interpolate(mode="nearest", size=n_out) (ATen
    nearest_neighbor_compute_source_index: min(int(floorf(dst * scale)), n_in - 1) with the fp32 scale n_in / n_out)."""
    import numpy as np

    scale = np.float32(n_in) / np.float32(n_out)
    idx = np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64)
    return torch.from_numpy(np.minimum(idx, n_in - 1).astype(np.int32))


def upsample_add(a: torch.Tensor, b: torch.Tensor, ymap: torch.Tensor, xmap: torch.Tensor) -> torch.Tensor:
    """a [n,h,w,c] + nearest-upsampled b [n,hb,wb,c]; ymap [h] / xmap [w] int32 device tensors (nearest_source_index)."""
    n, h, w, c = a.shape
    assert a.is_contiguous() and b.is_contiguous() and b.shape[0] == n and b.shape[3] == c and a.dtype == b.dtype
    assert ymap.dtype == torch.int32 and xmap.dtype == torch.int32 and ymap.numel() == h and xmap.numel() == w
    out = torch.empty_like(a)
    check(_L(a).avcer_upsample_add(a.data_ptr(), b.data_ptr(), n, h, w, b.shape[1], b.shape[2], c, ymap.data_ptr(), xmap.data_ptr(),
                                   out.data_ptr(), dtype_code(a.dtype), _stream()))
    return out


def det_decode(heads: Sequence[torch.Tensor], n: int, height: int, width: int) -> torch.Tensor:
    """heads: three fp32 [n * fh_k * fw_k, pitch] matrices (strides 8, 16, 32) -> dets [n, P, 15] fp32."""
    fhw = [((height + s - 1) // s, (width + s - 1) // s) for s in (8, 16, 32)]
    pitch = heads[0].stride(0)
    for hk, (fh, fw) in zip(heads, fhw):
        assert hk.dtype == torch.float32 and hk.shape[0] == n * fh * fw and hk.stride(0) == pitch and hk.stride(1) == 1
    P = 2 * sum(fh * fw for fh, fw in fhw)
    dets = torch.empty((n, P, 15), device=heads[0].device, dtype=torch.float32)
    check(_lib.load().avcer_det_decode(heads[0].data_ptr(), heads[1].data_ptr(), heads[2].data_ptr(), pitch, n, height, width,
                                       dets.data_ptr(), _stream()))
    return dets


def avgpool(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty((n, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_avgpool(x.data_ptr(), n, h * w, c, y.data_ptr(), dtype_code(x.dtype), _stream()))
    return y


def small_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], softmax: bool = False) -> torch.Tensor:
    n, k = x.shape
    m = w.shape[0]
    assert x.stride(1) == 1 and w.dtype == torch.float32 and w.is_contiguous()
    y = torch.empty((n, m), device=x.device, dtype=torch.float32)
    check(_L(x).avcer_small_linear(x.data_ptr(), n, k, x.stride(0), w.data_ptr(), _ptr(b), m, int(softmax), y.data_ptr(),
                                         dtype_code(x.dtype), _stream()))
    return y


def lstm_cell(xproj: Optional[torch.Tensor], xidx: Optional[torch.Tensor], hproj: Optional[torch.Tensor],
              c: torch.Tensor, h_out: torch.Tensor, hidden: int, first: bool, split: bool = False,
              h_f32: Optional[torch.Tensor] = None) -> None:
    """split: h_out is the [n, 3*hidden] bf16x3 operand [hi | lo | hi]; h_f32: optional fp32 copy of h [n, hidden]."""
    n = c.shape[0]
    check(_L(h_out).avcer_lstm_cell(_ptr(xproj), _ptr(xidx), _ptr(hproj), c.data_ptr(), h_out.data_ptr(),
                                      h_out.stride(0), n, hidden, int(first), int(split), _ptr(h_f32), dtype_code(h_out.dtype),
                                      _stream()))


def gru_cell(xg: torch.Tensor, x_row0: int, x_row_stride: int, hg: torch.Tensor, h_state: torch.Tensor, h_out: torch.Tensor,
             hidden: int, split: bool = False, y: Optional[torch.Tensor] = None, y_row0: int = 0, y_row_stride: int = 1) -> None:
    """One GRU time step for n = h_state.shape[0] sequences (see avcer_gru_cell)."""
    n = h_state.shape[0]
    assert xg.dtype == torch.float32 and hg.dtype == torch.float32 and h_state.dtype == torch.float32
    assert xg.is_contiguous() and hg.is_contiguous() and h_state.is_contiguous() and (y is None or y.is_contiguous())
    check(_L(h_out).avcer_gru_cell(xg.data_ptr(), x_row0, x_row_stride, hg.data_ptr(), h_state.data_ptr(), h_out.data_ptr(),
                                     h_out.stride(0), int(split), _ptr(y), y_row0, y_row_stride, n, hidden, dtype_code(h_out.dtype),
                                     _stream()))


def split_bf16x3(x: torch.Tensor, out: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """x fp32 [rows, k] -> 16-bit [rows, 3k] = [hi | lo | hi] (operand of a bf16x3 / fp16x3 contraction)."""
    _cuda(x, "x")
    rows, k = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    if out is None:
        out = torch.empty((rows, 3 * k), device=x.device, dtype=dtype)
    check(_L(out).avcer_split_bf16x3(x.data_ptr(), rows, k, x.stride(0), out.data_ptr(), out.stride(0), _stream()))
    return out


PAD_MODES = {"mean": 0, "constant": 1, "repeat": 2}


def sinc_resample_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Polyphase filter bank of torchaudio.transforms.Resample with its defaults (Hann-windowed sinc, computed in
    float64 and rounded to float32), the transform the reference applies at src/data/utils.py:53-55.
    Returns (bank [nnew, 2*width + orig] float32, orig, nnew, width) with the rates divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, nnew = int(orig_freq) // g, int(new_freq) // g
    cutoff = min(orig, nnew) * rolloff
    width = math.ceil(lowpass_filter_width * orig / cutoff)
    taps = np.arange(-width, width + orig, dtype=np.float64) / orig                  # tap position, in input periods
    phase = -np.arange(nnew, dtype=np.float64) / nnew                                # one filter per output phase
    t = np.clip((phase[:, None] + taps[None, :]) * cutoff, -lowpass_filter_width, lowpass_filter_width)
    hann = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    sinc = np.ones_like(t)
    nz = t != 0
    sinc[nz] = np.sin(t[nz]) / t[nz]
    return (sinc * hann * (cutoff / orig)).astype(np.float32), orig, nnew, width


def pcm16_to_mono(pcm: torch.Tensor, sr: int, sampling_rate: int) -> torch.Tensor:
    """pcm: device int16 [n, channels] (interleaved .wav payload) -> fp32 mono at `sampling_rate` (data/utils.py:49-60)."""
    _cuda(pcm, "pcm")
    if pcm.dtype != torch.int16 or pcm.dim() != 2 or not pcm.is_contiguous():
        raise ValueError("pcm16_to_mono expects a contiguous int16 [n, channels] tensor")
    n, ch = pcm.shape
    if sr == sampling_rate:
        out = torch.empty(n, device=pcm.device, dtype=torch.float32)
        check(_lib.load().avcer_pcm16_resample(pcm.data_ptr(), n, ch, None, 1, 1, 0, out.data_ptr(), n, _stream()))
        return out
    bank, orig, nnew, width = sinc_resample_bank(sr, sampling_rate)
    bank_t = torch.from_numpy(np.ascontiguousarray(bank.T)).to(pcm.device)
    n_out = -(-nnew * n // orig)
    out = torch.empty(n_out, device=pcm.device, dtype=torch.float32)
    check(_lib.load().avcer_pcm16_resample(pcm.data_ptr(), n, ch, bank_t.data_ptr(), orig, nnew, width, out.data_ptr(), n_out,
                                           _stream()))
    return out


def audio_normalize_windows(wav: torch.Tensor, starts: torch.Tensor, win: int, pad_mode: str,
                            ends: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wav: device fp32 buffer; chunk i = wav[starts[i]:ends[i]] (ends default: min(start+win, len(wav)))."""
    _cuda(wav, "wav")
    n_win = starts.numel()
    if ends is None:
        ends = torch.clamp(starts + win, max=wav.numel())
    if out is None:
        out = torch.empty((n_win, win), device=wav.device, dtype=torch.float32)
    check(_lib.load().avcer_audio_normalize_windows(wav.data_ptr(), starts.data_ptr(), ends.data_ptr(), n_win, win,
                                                    PAD_MODES[pad_mode], out.data_ptr(), _stream()))
    return out


def w2v_conv0_ln_gelu(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, g: torch.Tensor, be: torch.Tensor,
                      y: torch.Tensor) -> torch.Tensor:
    n, t_in = x.shape
    check(_L(y).avcer_w2v_conv0_ln_gelu(x.data_ptr(), n, t_in, w.data_ptr(), b.data_ptr(), g.data_ptr(),
                                              be.data_ptr(), y.data_ptr(), y.shape[1], dtype_code(y.dtype), _stream()))
    return y


def w2v_conv0_tc(x: torch.Tensor, w_packed: torch.Tensor, g: torch.Tensor, be: torch.Tensor, y: torch.Tensor,
                 eps: float = 1e-5) -> torch.Tensor:
    """Same contract as w2v_conv0_ln_gelu on the tensor cores (16-bit libraries only; w_packed = weights.pack_conv0_tc)."""
    n, t_in = x.shape
    assert y.dtype in (torch.bfloat16, torch.float16) and w_packed.dtype == y.dtype and y.shape[2] == 512 and y.stride(1) == 512
    with _Timed("w2v_conv0_tc", 0.0):
        check(_L(y).avcer_w2v_conv0_tc(x.data_ptr(), n, t_in, w_packed.data_ptr(), g.data_ptr(), be.data_ptr(), eps,
                                       y.data_ptr(), y.stride(0) // 512, _stream()))
    return y


def layernorm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float, *, act: int = ACT_NONE,
              add: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [rows, c] with row pitch x.stride(0)."""
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_layernorm(x.data_ptr(), rows, c, x.stride(0), _ptr(add), 0 if add is None else add.shape[0],
                                      g.data_ptr(), b.data_ptr(), eps, act, out.data_ptr(), out.stride(0),
                                      dtype_code(x.dtype), _stream()))
    return out


def add_rows(x: torch.Tensor, add: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, c = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(_L(x).avcer_add_rows(x.data_ptr(), rows, c, add.data_ptr(), add.shape[0], out.data_ptr(),
                                     dtype_code(x.dtype), _stream()))
    return out


def attention(qkv: torch.Tensor, n: int, t: int, heads: int, dh: int, scale: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv: [n*t, 3*heads*dh] packed (q | k | v); returns [n*t, heads*dh]."""
    assert qkv.is_contiguous()
    if out is None:
        out = torch.empty((n * t, heads * dh), device=qkv.device, dtype=qkv.dtype)
    check(_L(qkv).avcer_attention(qkv.data_ptr(), n, t, heads, dh, scale, out.data_ptr(), dtype_code(qkv.dtype), _stream()))
    return out


def maxpool1d5_relu(x: torch.Tensor) -> torch.Tensor:
    n, t, c = x.shape
    y = torch.empty((n, t // 5, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_maxpool1d5_relu(x.data_ptr(), n, t, c, y.data_ptr(), dtype_code(x.dtype), _stream()))
    return y


def avgpool1d_relu(x: torch.Tensor) -> torch.Tensor:
    n, t, c = x.shape
    y = torch.empty((n, c), device=x.device, dtype=x.dtype)
    check(_L(x).avcer_avgpool1d_relu(x.data_ptr(), n, t, c, y.data_ptr(), dtype_code(x.dtype), _stream()))
    return y


def cast(x: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    y = torch.empty(x.shape, device=x.device, dtype=dt)
    check(_L(x, y).avcer_cast(x.data_ptr(), x.numel(), dtype_code(x.dtype), y.data_ptr(), dtype_code(dt), _stream()))
    return y
