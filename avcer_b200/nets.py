"""Device-side forward passes of the three single-modality networks, expressed as sequences of
libavcer_b200 kernel launches (host orchestration only -- no torch math on the data path).

  VSNet : ResNet-50 static visual model          (reference src/architectures/video.py:93-166)
  VDNet : 2-layer LSTM over sliding windows      (reference src/architectures/video.py:169-185)
  ANet  : wav2vec2-large-robust + 2 TL + head    (reference src/architectures/audio_8_cl.py:131-190)

precision "bf16": bf16 storage, tcgen05 tensor-core contractions with fp32 accumulation (libavcer_b200.so).
precision "fp16": the same kernels built with IEEE half storage (libavcer_b200_fp16.so): same speed, 11 mantissa bits.
precision "fp32": fp32 storage, SIMT contractions (the tolerance-check mode of the north star).
"""
from __future__ import annotations

import gc
import math
from typing import List, Dict, Optional, Tuple

import torch

from . import ops, weights
from . import _lib
from ._lib import require_device

PAD_H, PAD = ops.PAD_H, ops.PAD_W      # 232 rows x 240-pixel pitch of the zero-bordered stem input
_BIG = 1 << 30


class GraphedForward:
    """CUDA-graph cache around a fixed-shape forward: the ~60 (VS) / ~130 (A) kernel launches of one
    batch are captured once per (batch size, static input buffer) and replayed with a single launch."""

    def __init__(self, fn, sm_limit: int = 0):
        self.fn = fn
        self.cache = {}
        self.sm_limit = sm_limit        # grid share of the persistent kernels baked into the captured launches
        self.max_entries = 24

    def __call__(self, *static_inputs: torch.Tensor):
        key = tuple((t.data_ptr(), tuple(t.shape)) for t in static_inputs)
        entry = self.cache.get(key)
        if entry is None:
            if len(self.cache) >= self.max_entries:          # bounded: ragged workloads would otherwise keep every shape's graph
                self.cache.pop(next(iter(self.cache)))
            prof, ops.PROFILE = ops.PROFILE, None
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side), ops.sm_limit(self.sm_limit):   # warm-up outside capture (lazy attribute sets, allocator)
                self.fn(*static_inputs)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            n0 = ops.STATS["launches"]
            graph = torch.cuda.CUDAGraph()
            # No garbage collection while the stream is capturing: an unreachable Engine (its CUDA graphs sit in reference
            # cycles) finalised in the middle of a capture destroys its graphs -- an operation that is not permitted
            # during stream capture and invalidates the one in progress.  Collect first, then hold the collector off.
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph), ops.sm_limit(self.sm_limit):
                    outs = self.fn(*static_inputs)
            finally:
                if gc_was_on:
                    gc.enable()
            entry = (graph, outs, ops.STATS["launches"] - n0)
            ops.STATS["launches"] = n0
            self.cache[key] = entry
            ops.PROFILE = prof
        graph, outs, n_launch = entry
        graph.replay()
        ops.STATS["launches"] += n_launch
        return outs


def _dtype(precision: str) -> torch.dtype:
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp16":
        return torch.float16
    if precision == "fp32":
        return torch.float32
    raise ValueError(f"precision must be 'bf16', 'fp16' or 'fp32', got {precision!r}")


def _tc(dtype: torch.dtype) -> bool:
    """16-bit storage = the tensor-core path (bf16 or fp16 build of the library)."""
    return dtype in (torch.bfloat16, torch.float16)


class VSNet:
    """Static visual ResNet-50.  forward(x) with x = zero-bordered NHWC4 crops [n,232,240,4]
    (layout 1/2 of avcer_preprocess_u8) -> (probabilities [n,7] fp32, relu(fc1) features [n,512] fp32)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], precision: str = "bf16", device: str = "cuda:0"):
        require_device()
        self.dtype = _dtype(precision)
        _lib.load("fp16" if self.dtype == torch.float16 else "bf16")       # the matching build of the library, loaded up front
        self.precision = precision
        self.device = torch.device(device)
        self.w = weights.pack_vs(state_dict, self.device, self.dtype)
        self.fused_stem = True          # bf16: stem + max-pool in one kernel (False: two kernels, same bits)
        self.fused_shortcut = True      # bf16: projection shortcuts folded into conv3 (K-concatenated GEMM, first block of a stage)
        self.alternate = True           # bf16: consecutive contractions walk their tiles in opposite directions (L2 reuse, _rev)
        self._dir = False
        self.sampled_tail = 2           # bf16: the last conv3 (1) and conv2 (2) of layer1-3 only at the pixels the next stage samples

    @property
    def input_layout(self) -> int:
        return 1 if _tc(self.dtype) else 2

    def alloc_input(self, n: int) -> torch.Tensor:
        """Zero-bordered input buffer; the border is written once here, K1 only fills the interior."""
        return torch.zeros((n, PAD_H, PAD, 4), device=self.device, dtype=self.dtype)

    def stem(self, x: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        st = self.w["stem"]
        y = torch.empty((n, 112, 112, 64), device=self.device, dtype=self.dtype)
        if _tc(self.dtype):
            # Strip mode: one padded image row (240 px x 4 ch = 1920 B = 15 x 128 B) is fetched once per (oy, ky);
            # output ox reads the 32 elements starting 16 B * ox into it (overlap expressed in the MMA descriptor).
            ops.contract(a=x, a_dim=(64, PAD * 4 // 64, 112, n, 7), a_stride=(1, 64, 2 * PAD * 4, PAD_H * PAD * 4, PAD * 4),
                         wt=st.wt, bias=st.bias, out=y, out_stride=(64, 112 * 64, 112 * 112 * 64), W=112, H=112, NB=n,
                         cin=32, cout=64, taps_w=1, taps_h=7, tap_h_in_dim4=True, act=ops.ACT_RELU, algo_k=147,
                         a_strip=True, wt_packed=self.w["stem_packed"])
        else:
            # A view: (8 pixels x 4 ch = 32, ox stride 2 px, oy stride 2 rows, n, ky stride 1 row)
            ops.contract(a=x, a_dim=(32, 112, 112, n, 7), a_stride=(1, 8, 2 * PAD * 4, PAD_H * PAD * 4, PAD * 4),
                         wt=st.wt, bias=st.bias, out=y, out_stride=(64, 112 * 64, 112 * 112 * 64), W=112, H=112, NB=n,
                         cin=32, cout=64, taps_w=1, taps_h=7, tap_h_in_dim4=True, act=ops.ACT_RELU,
                         algo_k=147)          # 7x7x3 real taps; the padded pixel/channel carry zero weights
        return y

    def _rev(self) -> bool:
        """Tile direction of the next contraction.  Every layer1 / layer2 tensor of a full batch is larger than the 126 MB L2
        (256 crops: 99 - 396 MB), so a consumer that walks its tiles in the producer's order finds its first rows evicted;
        walking them backwards it starts on what the producer wrote last.  Directions simply alternate from launch to launch
        (same tiles, same arithmetic: bit-identical results)."""
        if not (self.alternate and _tc(self.dtype)):
            return False
        self._dir = not self._dir
        return self._dir

    def _conv(self, x: torch.Tensor, pc: weights.PackedConv, act: int, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        pad = (pc.k - 1) // 2
        return ops.conv2d_nhwc(x, pc.wt, pc.bias, kh=pc.k, kw=pc.k, stride=pc.stride, pad_h=pad, pad_w=pad,
                               residual=residual, act=act, reverse=self._rev())

    @property
    def k1_fused(self) -> bool:
        """bf16 + fused stem: packed 224x224 uint8 crops can enter the network directly (forward_u8)."""
        return _tc(self.dtype) and self.fused_stem and self.fused_shortcut and "conv3_ds" in self.w["blocks"][0]

    def stem_u8(self, crops_u8: torch.Tensor, cat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K1 + stem + max-pool in one kernel (avcer_stem_pool_u8): uint8 [n,224,224,3] BGR crops -> the left 64 channels of
        the [n,55,55,128] matrix layer1.0 works on.  Bit-identical to preprocess(layout 1) + forward's own stem."""
        assert self.k1_fused
        if cat is None:
            cat = torch.empty((crops_u8.shape[0], 55, 55, 128), device=self.device, dtype=self.dtype)
        ops.stem_pool_u8(crops_u8, self.w["stem_packed"], self.w["stem"].bias, out=cat[..., :64])
        return cat

    def forward_u8(self, crops_u8: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.body(self.stem_u8(crops_u8))

    def forward(self, x: torch.Tensor, taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        assert x.shape[1:] == (PAD_H, PAD, 4) and x.dtype == self.dtype
        cat = None
        if taps is None and _tc(self.dtype) and self.fused_stem:
            if self.fused_shortcut and "conv3_ds" in self.w["blocks"][0]:
                cat = torch.empty((x.shape[0], 55, 55, 128), device=self.device, dtype=self.dtype)
                y = ops.stem_pool(x, self.w["stem_packed"], self.w["stem"].bias, out=cat[..., :64])
            else:
                y = ops.stem_pool(x, self.w["stem_packed"], self.w["stem"].bias)      # stem activation stays on chip
        else:
            y = self.stem(x)
            if taps is not None:
                taps["stem"] = y
            y = ops.maxpool3x3s2(y)
            if taps is not None:
                taps["pool"] = y
        return self.body(cat if cat is not None else y, taps, cat is not None)

    def body(self, y: torch.Tensor, taps: Optional[dict] = None, is_cat: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """layer1 .. fc2 from the pooled stem output: `y` is either [n,55,55,64] or (is_cat) the [n,55,55,128] matrix whose
        left 64 channels hold it (layer1.0's K-concatenated layout)."""
        cat = y if is_cat and y.shape[-1] == 128 else None
        y, _, _ = self._blocks(y, cat, None, 0, len(self.w["blocks"]), taps)
        return self._tail(y)

    def _blocks(self, y, cat, cat4, lo: int, hi: int, taps: Optional[dict] = None, tail_out: Optional[torch.Tensor] = None):
        if lo == 0:
            self._dir = False           # the stem walks forward; the first contraction after it walks backwards
        """Bottleneck blocks [lo, hi).  State between blocks: `y` (block input, NHWC) or, after a sampled stage tail, the
        K-concatenated matrix `cat` / `cat4` whose left columns hold the stride-2 sampled input of the coming block.
        `tail_out`: where the sampled tail of block hi-1 puts that matrix (rows of a batch slice of a larger one)."""
        blocks = self.w["blocks"]
        fuse = taps is None and _tc(self.dtype) and self.fused_shortcut
        sampled = cat4 is not None      # `cat[:, :cin]` already holds the stride-2 sampled input of the coming block
        for bi in range(lo, hi):
            blk = blocks[bi]
            if bi == 0 and cat is not None:
                # layer1.0: block input and conv2 output share one [n*55*55, 128] matrix, so conv3 and the projection
                # shortcut are a single K = 128 GEMM (no shortcut tensor, no residual read)
                m = cat.shape[0] * 55 * 55
                c1, c2, c3 = blk["conv1"], blk["conv2"], blk["conv3_ds"]
                t = ops.linear(cat.view(m, 128)[:, :64], c1.wt, c1.bias, act=ops.ACT_RELU, reverse=self._rev()).view(-1, 55, 55, 64)
                ops.conv2d_nhwc(t, c2.wt, c2.bias, kh=3, kw=3, pad_h=1, pad_w=1, act=ops.ACT_RELU, out=cat[..., 64:], reverse=self._rev())
                y = ops.linear(cat.view(m, 128), c3.wt, c3.bias, act=ops.ACT_RELU, reverse=self._rev()).view(-1, 55, 55, 256)
                continue
            if fuse and "conv3_ds" in blk and blk["conv1"].stride == 2:
                # layer2-4 block 0: the stride-2 sampling of the block input is materialised once, as the first Cin
                # columns of the matrix whose remaining columns conv2 fills; conv1 is a plain row GEMM over it and
                # conv3 + projection shortcut one K = Cin + planes GEMM
                c1, c2, c3 = blk["conv1"], blk["conv2"], blk["conv3_ds"]
                if sampled:                                  # the previous block wrote the sampled pixels itself
                    nb, ho, wo, cin = cat4.shape[0], cat4.shape[1], cat4.shape[2], c1.cin
                    m = nb * ho * wo
                    sampled = False
                else:
                    nb, hh, ww, cin = y.shape
                    ho, wo = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                    m = nb * ho * wo
                    cat = torch.empty((m, cin + c1.cout), device=self.device, dtype=self.dtype)
                    ops.subsample_rows(y, 2, cat[:, :cin])
                t = ops.linear(cat[:, :cin], c1.wt, c1.bias, act=ops.ACT_RELU, reverse=self._rev()).view(nb, ho, wo, c1.cout)
                ops.conv2d_nhwc(t, c2.wt, c2.bias, kh=3, kw=3, pad_h=1, pad_w=1, act=ops.ACT_RELU,
                                out=cat.view(nb, ho, wo, cin + c1.cout)[..., cin:], reverse=self._rev())
                y = ops.linear(cat, c3.wt, c3.bias, act=ops.ACT_RELU, reverse=self._rev()).view(nb, ho, wo, c3.cout)
                continue
            identity = self._conv(y, blk["ds"], ops.ACT_NONE) if "ds" in blk else y
            t = self._conv(y, blk["conv1"], ops.ACT_RELU)
            nxt = blocks[bi + 1] if bi + 1 < len(blocks) else None
            if fuse and self.sampled_tail and "ds" not in blk and nxt is not None and "conv3_ds" in nxt and nxt["conv1"].stride == 2:
                # Last block of a stage: the next stage reads its output only at every second pixel (stride-2 conv1 and
                # projection shortcut, video.py:13-15, 141-148).  conv3 is pointwise, so conv3 + residual + ReLU are
                # computed at those pixels only (a quarter of the rows) and land directly in the left columns of the next
                # block's K-concatenated matrix; the full-resolution stage output is never formed.  conv2's output is
                # consumed by that conv3 only, so (sampled_tail >= 2) the 3x3 conv too is evaluated at every second pixel:
                # a stride-2 "same" conv over the full-resolution conv1 output (TMA traversal stride 2).
                c2, c3 = blk["conv2"], blk["conv3"]
                c3_stride = 2
                if self.sampled_tail >= 2:
                    t = ops.conv2d_nhwc(t, c2.wt, c2.bias, kh=3, kw=3, stride=2, pad_h=1, pad_w=1, act=ops.ACT_RELU, reverse=self._rev())
                    c3_stride = 1
                else:
                    t = self._conv(t, c2, ops.ACT_RELU)
                nb, hh, ww, _ = t.shape
                ho, wo = (hh - 1) // c3_stride + 1, (ww - 1) // c3_stride + 1
                if tail_out is not None and bi == hi - 1:
                    cat = tail_out
                    assert cat.shape == (nb * ho * wo, c3.cout + nxt["conv1"].cout)
                else:
                    cat = torch.empty((nb * ho * wo, c3.cout + nxt["conv1"].cout), device=self.device, dtype=self.dtype)
                cat4 = cat.view(nb, ho, wo, c3.cout + nxt["conv1"].cout)
                ops.conv2d_nhwc(t, c3.wt, c3.bias, kh=1, kw=1, stride=c3_stride, residual=identity, residual_stride=2,
                                act=ops.ACT_RELU, out=cat4[..., :c3.cout], reverse=self._rev())
                sampled = True
                y = None
                continue
            t = self._conv(t, blk["conv2"], ops.ACT_RELU)
            y = self._conv(t, blk["conv3"], ops.ACT_RELU, residual=identity)
            if taps is not None:
                taps[f"block{bi}"] = y
        return y, cat, (cat4 if sampled else None)

    def _tail(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        pooled = ops.avgpool(y)
        fc1 = self.w["fc1"]
        # relu(fc1) in fp32 (direct-store epilogue): fc2 input and, split into bf16x3, the LSTM input (VDNet.forward)
        feat = ops.linear(pooled, fc1.wt, fc1.bias, act=ops.ACT_RELU, out_dtype=torch.float32)
        probs = ops.small_linear(feat, self.w["fc2_w"], self.w["fc2_b"], softmax=True)
        return probs, feat


class RetinaFaceNet:
    """RetinaFace-ResNet50 face detector (retina_face.py:48-115 of the reference's ibug package):
    detect(frames) with frames = uint8 [n,H,W,3] video frames on the device -> decoded rows [n, P, 15] fp32
    (x1, y1, x2, y2, score, 5 landmarks; priors in the reference's order).  Any frame size; a batch shares one size."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], precision: str = "bf16", device: str = "cuda:0"):
        require_device()
        self.dtype = _dtype(precision)
        _lib.load("fp16" if self.dtype == torch.float16 else "bf16")
        self.device = torch.device(device)
        self.w = weights.pack_retinaface(state_dict, self.device, self.dtype)
        self.stem_tc = True            # 16-bit modes: stem as strip-mode tcgen05 contractions (False: the fp32 direct-conv kernel)
        self._maps: Dict[tuple, torch.Tensor] = {}

    def _conv(self, x, pc: weights.PackedConv, act: int, residual=None, out=None) -> torch.Tensor:
        pad = (pc.k - 1) // 2
        if pc.k == 3 and pc.stride == 2 and not _tc(self.dtype):
            # fp32 mode has no strided k x k implicit GEMM: stride-1 conv, then every second pixel (the same values)
            full = ops.conv2d_nhwc(x, pc.wt, pc.bias, kh=3, kw=3, stride=1, pad_h=1, pad_w=1, act=act)
            n, h, w, c = full.shape
            y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), device=self.device, dtype=self.dtype)
            ops.subsample_rows(full, 2, y.view(-1, c))
            return y
        return ops.conv2d_nhwc(x, pc.wt, pc.bias, kh=pc.k, kw=pc.k, stride=pc.stride, pad_h=pad, pad_w=pad, residual=residual,
                               act=act, out=out)

    def _map(self, n_in: int, n_out: int) -> torch.Tensor:
        key = (n_in, n_out)
        if key not in self._maps:
            self._maps[key] = ops.nearest_source_index(n_in, n_out).to(self.device)
        return self._maps[key]

    def _merge(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return ops.upsample_add(a, b, self._map(b.shape[1], a.shape[1]), self._map(b.shape[2], a.shape[2]))

    def _ssh(self, x: torch.Tensor, w: dict) -> torch.Tensor:
        """retina_face_net.py:43-62; relu(cat(...)) = every branch's last conv with a ReLU epilogue writing its channel slice."""
        n, h, ww, _ = x.shape
        out = torch.empty((n, h, ww, 256), device=self.device, dtype=self.dtype)
        self._conv(x, w["conv3X3"], ops.ACT_RELU, out=out[..., :128])
        c5_1 = self._conv(x, w["conv5X5_1"], ops.ACT_RELU)
        self._conv(c5_1, w["conv5X5_2"], ops.ACT_RELU, out=out[..., 128:192])
        c7_2 = self._conv(c5_1, w["conv7X7_2"], ops.ACT_RELU)
        self._conv(c7_2, w["conv7x7_3"], ops.ACT_RELU, out=out[..., 192:])
        return out

    def heads(self, frames: torch.Tensor, rgb: bool = False, taps: Optional[dict] = None) -> List[torch.Tensor]:
        """-> per pyramid level (strides 8, 16, 32) the fp32 [n*fh*fw, 64] head matrix: columns 0-3 class logits
        (anchor 0: background, face; anchor 1), 4-11 box, 12-31 landmark regressions."""
        w = self.w
        if "stem_packed" in w and self.stem_tc:
            y = ops.det_stem_tc(frames, w["stem_wt"], w["stem_packed"], w["stem_b"], self.dtype, rgb)
        else:
            y = ops.det_stem(frames, w["stem_w"], w["stem_b"], self.dtype, rgb)
        if taps is not None:
            taps["stem"] = y
        y = ops.maxpool3x3s2p1(y)
        if taps is not None:
            taps["pool"] = y
        feats = []
        for blk in w["blocks"]:
            idn = self._conv(y, blk["ds"], ops.ACT_NONE) if "ds" in blk else y
            t = self._conv(y, blk["conv1"], ops.ACT_RELU)
            t = self._conv(t, blk["conv2"], ops.ACT_RELU)
            y = self._conv(t, blk["conv3"], ops.ACT_RELU, residual=idn)
            if blk["last_of_layer"]:
                if taps is not None:
                    taps[f"layer{blk['layer']}"] = y
                if blk["layer"] >= 2:
                    feats.append(y)
        o1, o2, o3 = (self._conv(f, pc, ops.ACT_RELU) for f, pc in zip(feats, w["fpn_out"]))
        o2 = self._conv(self._merge(o2, o3), w["fpn_merge"][1], ops.ACT_RELU)
        o1 = self._conv(self._merge(o1, o2), w["fpn_merge"][0], ops.ACT_RELU)
        out = []
        for i, (o, sw, hw) in enumerate(zip((o1, o2, o3), w["ssh"], w["heads"])):
            f = self._ssh(o, sw)
            if taps is not None:
                taps[f"fpn{i + 1}"], taps[f"ssh{i + 1}"] = o, f
            out.append(ops.linear(f.view(-1, 256), hw.wt, hw.bias, out_dtype=torch.float32))
        return out

    def detect(self, frames: torch.Tensor, rgb: bool = False) -> torch.Tensor:
        n, h, w, _ = frames.shape
        return ops.det_decode(self.heads(frames, rgb), n, h, w)


class VDNet:
    """Dynamic visual LSTM over sliding windows of VS features.  All windows are advanced together,
    one recurrent GEMM per step; the layer-1 input projection is computed once per unique feature."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], precision: str = "bf16", device: str = "cuda:0"):
        require_device()
        self.dtype = _dtype(precision)
        _lib.load("fp16" if self.dtype == torch.float16 else "bf16")       # the matching build of the library, loaded up front
        self.device = torch.device(device)
        self.w = weights.pack_vd(state_dict, self.device, self.dtype)

    def forward(self, feats: torch.Tensor, windows_t: torch.Tensor) -> torch.Tensor:
        """feats: [U,512] fp32 relu(fc1) features; windows_t: int32 [10, M] positions into feats
        (time-major).  Returns VD logits [M,7] fp32.
        bf16 mode runs every contraction as bf16x3 (activations [hi | lo | hi] x weights [hi | hi | lo], fp32
        accumulation): the 20 chained recurrent GEMMs keep 16 mantissa bits at 3x the (negligible) tensor work."""
        steps, m = windows_t.shape
        if m == 0:
            return torch.empty((0, 7), device=self.device, dtype=torch.float32)
        assert feats.dtype == torch.float32
        w = self.w
        split = w["split"]
        s = 3 if split else 1
        x = ops.split_bf16x3(feats, dtype=self.dtype) if split else feats
        xproj = ops.linear(x, w["w_ih1"], w["b1"], out_dtype=torch.float32)                 # [U, 2048]
        hcat = torch.zeros((m, s * 768), device=self.device, dtype=self.dtype)               # [h1_t | h2_{t-1}] (each x3 when split)
        c1 = torch.empty((m, 512), device=self.device, dtype=torch.float32)
        c2 = torch.empty((m, 256), device=self.device, dtype=torch.float32)
        h1, h2 = hcat[:, :s * 512], hcat[:, s * 512:]
        h2_f32 = torch.empty((m, 256), device=self.device, dtype=torch.float32)
        hproj = torch.empty((m, 2048), device=self.device, dtype=torch.float32)
        g2 = torch.empty((m, 1024), device=self.device, dtype=torch.float32)
        for t in range(steps):
            if t > 0:
                ops.linear(h1, w["w_hh1"], None, out=hproj)
            ops.lstm_cell(xproj, windows_t[t], hproj if t > 0 else None, c1, h1, 512, first=(t == 0), split=split)
            ops.linear(hcat, w["w_cat2"], w["b2"], out=g2)
            ops.lstm_cell(None, None, g2, c2, h2, 256, first=(t == 0), split=split, h_f32=h2_f32 if t == steps - 1 else None)
        return ops.small_linear(h2_f32, w["fc_w"], w["fc_b"], softmax=False)


W2V_LENGTHS = (12799, 6399, 3199, 1599, 799, 399, 199)     # conv output lengths for 64000 samples
W2V_KERNELS = (10, 3, 3, 3, 3, 2, 2)


class ANet:
    """Audio emotion network over normalised 4 s windows: x [B,64000] fp32 -> logits [B,ncls] fp32."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], precision: str = "bf16", device: str = "cuda:0"):
        require_device()
        self.dtype = _dtype(precision)
        _lib.load("fp16" if self.dtype == torch.float16 else "bf16")       # the matching build of the library, loaded up front
        self.device = torch.device(device)
        self.w = weights.pack_audio(state_dict, self.device, self.dtype)
        self.num_classes = self.w["num_classes"]
        self.conv0_tc = True       # 16-bit modes: feature-extractor layer 0 on the tensor cores (avcer_w2v_conv0_tc)

    def _conv1d_s2(self, x: torch.Tensor, t_in: int, k: int, wt, bias) -> torch.Tensor:
        """[B, t_in, 512] -> [B, t_out, 512], stride 2, no padding: the k taps of one output step are
        contiguous in memory, so the layer is a plain GEMM over an overlapping strided view."""
        b = x.shape[0]
        t_out = (t_in - k) // 2 + 1
        y = torch.empty((b, t_out, 512), device=self.device, dtype=self.dtype)
        ops.contract(a=x, a_dim=(k * 512, t_out, 1, b, 1), a_stride=(1, 2 * 512, _BIG, t_in * 512, _BIG), wt=wt, bias=bias,
                     out=y, out_stride=(512, 0, t_out * 512), W=t_out, H=1, NB=b, cin=k * 512, cout=512)
        return y

    def _encoder_layer(self, h: torch.Tensor, L: dict, b: int, t: int) -> torch.Tensor:
        a = ops.layernorm(h, *L["ln1"], 1e-5)
        qkv = ops.linear(a, L["wqkv"], L["bqkv"])
        att = ops.attention(qkv, b, t, 16, 64, 0.125)
        h = ops.linear(att, L["wo"], L["bo"], residual=h)
        f = ops.layernorm(h, *L["ln2"], 1e-5)
        f = ops.linear(f, L["w1"], L["b1"], act=ops.ACT_GELU)
        return ops.linear(f, L["w2"], L["b2"], residual=h)

    def _transformer_layer(self, h: torch.Tensor, T: dict, b: int, t: int) -> torch.Tensor:
        heads = T["heads"]
        dh = 1024 // heads
        xp = ops.add_rows(h, T["pe"][:t])
        qkv = ops.linear(xp, T["wqkv"], None)
        att = ops.attention(qkv, b, t, heads, dh, 1.0 / math.sqrt(dh))
        y = ops.linear(att, T["wo"], None, residual=xp)
        y = ops.layernorm(y, *T["ln1"], 1e-5)
        f = ops.linear(y, T["w1"], T["b1"], act=ops.ACT_RELU)
        f = ops.linear(f, T["w2"], T["b2"], residual=y)
        return ops.layernorm(f, *T["ln2"], 1e-5)

    def forward(self, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        assert x.dtype == torch.float32 and x.shape[1] == 64000
        b = x.shape[0]
        w = self.w
        t0 = W2V_LENGTHS[0]
        h = torch.empty((b, t0, 512), device=self.device, dtype=self.dtype)
        if self.conv0_tc and "conv0_tc" in w:
            ops.w2v_conv0_tc(x, w["conv0_tc"], *w["conv_ln"][0], h)
        else:
            ops.w2v_conv0_ln_gelu(x, w["conv0_w"], w["conv0_b"], *w["conv_ln"][0], h)
        if taps is not None:
            taps["conv0"] = h
        t = t0
        for i in range(1, 7):
            wt, bias = w["convs"][i - 1]
            h = self._conv1d_s2(h, t, W2V_KERNELS[i], wt, bias)
            t = h.shape[1]
            g, be = w["conv_ln"][i]
            h2 = h.view(b * t, 512)
            ops.layernorm(h2, g, be, 1e-5, act=ops.ACT_GELU, out=h2)
        if taps is not None:
            taps["conv6"] = h
        rows = h.view(b * t, 512)
        rows = ops.layernorm(rows, *w["fp_ln"], 1e-5)
        h = ops.linear(rows, w["fp_w"], w["fp_b"])                                       # [b*t, 1024]
        if taps is not None:
            taps["proj"] = h
        # grouped positional conv (k=128, pad 64, 16 groups), last frame dropped, GELU, added to h
        h2 = torch.empty_like(h)
        ops.contract(a=h, a_dim=(1024, t, 1, b, 1), a_stride=(1, 1024, _BIG, t * 1024, _BIG), wt=w["pos_w"], bias=w["pos_b"],
                     out=h2, out_stride=(1024, 0, t * 1024), W=t, H=1, NB=b, cin=64, cout=1024, taps_w=128, off_w=-64,
                     group_cin_shift=64, residual=h, act=ops.ACT_GELU, res_after_act=True)
        h = h2
        if taps is not None:
            taps["posconv"] = h
        for li, L in enumerate(w["layers"]):
            h = self._encoder_layer(h, L, b, t)
            if taps is not None:
                taps[f"layer{li}"] = h
        h = ops.layernorm(h, *w["enc_ln"], 1e-5)
        if taps is not None:
            taps["w2v"] = h
        c = w["f_size"]
        if w["variant"] == "v1":
            h = self._gru(h, b, t)                       # [b*t, 256]
            if taps is not None:
                taps["gru"] = h
        else:
            h = self._transformer_layer(h, w["tl1"], b, t)
            h = self._transformer_layer(h, w["tl2"], b, t)
            if taps is not None:
                taps["tl2"] = h
        # time_downsample: Conv1d(k5,s3,d2)+BN -> MaxPool1d(5) -> ReLU -> Conv1d(k3)+BN -> mean -> ReLU
        t1 = (t - 2 * 4 - 1) // 3 + 1
        y = torch.empty((b, t1, c), device=self.device, dtype=self.dtype)
        ops.contract(a=h, a_dim=(c, t1, 1, b, 5), a_stride=(1, 3 * c, _BIG, t * c, 2 * c), wt=w["td0_w"],
                     bias=w["td0_b"], out=y, out_stride=(c, 0, t1 * c), W=t1, H=1, NB=b, cin=c, cout=c,
                     taps_w=1, taps_h=5, tap_h_in_dim4=True)
        y = ops.maxpool1d5_relu(y)
        t2 = y.shape[1]
        t3 = t2 - 2
        z = torch.empty((b, t3, c), device=self.device, dtype=self.dtype)
        ops.contract(a=y, a_dim=(3 * c, t3, 1, b, 1), a_stride=(1, c, _BIG, t2 * c, _BIG), wt=w["td4_w"],
                     bias=w["td4_b"], out=z, out_stride=(c, 0, t3 * c), W=t3, H=1, NB=b, cin=3 * c, cout=c)
        z = ops.avgpool1d_relu(z)
        return ops.small_linear(z, w["fd_w"], w["fd_b"], softmax=False)

    def _gru(self, x: torch.Tensor, b: int, t: int) -> torch.Tensor:
        """ExprModelV1's nn.GRU(1024 -> 256, 2 layers, batch_first) from zero state (audio_8_cl.py:23-29,63): per layer one
        contraction for the input projections of all time steps, then t recurrent steps of (h W_hh^T + b_hh, gate math).
        x: [b*t, Cin] rows in (window, time) order -> [b*t, 256]."""
        split = _tc(self.dtype)
        s = 3 if split else 1
        for L in self.w["gru"]:
            xg = ops.linear(x, L["w_ih"], L["b_ih"], out_dtype=torch.float32)                  # [b*t, 768]
            h_state = torch.zeros((b, 256), device=self.device, dtype=torch.float32)
            h_op = torch.zeros((b, s * 256), device=self.device, dtype=self.dtype)             # operand of the recurrent GEMM
            hg = torch.empty((b, 768), device=self.device, dtype=torch.float32)
            y = torch.empty((b * t, 256), device=self.device, dtype=self.dtype)
            for step in range(t):
                ops.linear(h_op, L["w_hh"], L["b_hh"], out=hg)
                ops.gru_cell(xg, step, t, hg, h_state, h_op, 256, split=split, y=y, y_row0=step, y_row_stride=t)
            x = y
        return x
