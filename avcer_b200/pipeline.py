"""Batched inference-and-fusion engine: the per-frame / per-window Python loops of the reference
(src/get_prob_video.py:91-180, src/get_prob_audio_8_cl.py:78-101, src/run.py:76-165) restated as
index plans (host, integer bookkeeping) + batched kernel launches (device).

Vocabulary follows the reference: a *clip* has N frames (face crops, some may be missing) and one
16 kHz waveform; VS rows are per-frame probabilities, VD rows per-frame logits carried forward from
the latest 10-slot window, A rows per-frame means of the logits of all 4 s windows covering the frame.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .nets import ANet, GraphedForward, VDNet, VSNet

VIDEO_ORDER = ["Neutral", "Happiness", "Sadness", "Surprise", "Fear", "Disgust", "Anger"]     # get_prob_video.py:56-64
AUDIO_ORDER = ["Neutral", "Anger", "Disgust", "Fear", "Happiness", "Sadness", "Surprise", "Other"]  # run.py:56-65
VIDEO_TO_AUDIO = [0, 6, 5, 4, 1, 2, 3]          # audio-order column j = video column VIDEO_TO_AUDIO[j]


# =========================================================================================== host index plans
def vd_step(fps: float) -> int:
    """get_prob_video.py:77 -- Python round() (half to even)."""
    return round((5 * fps) / 25)


@dataclass
class VideoPlan:
    samples: np.ndarray     # [M] frame indices feeding the LSTM
    windows: np.ndarray     # [M,10] positions into `samples`
    stat_src: np.ndarray    # [N] frame index whose VS row is shown, -1 = zero row
    dyn_src: np.ndarray     # [N] window index whose VD row is shown, -1 = zero row


def plan_video(exists: Sequence[bool], step: int) -> VideoPlan:
    """Closed-form equivalent of the frame loop of get_prob_video.py:91-178.

    * an existing frame i with i % step == 0 is a *sample*; any missing frame resets the 10-slot
      window (:169), so windows never reach across a gap; the first sample of a gap-free segment
      is repeated on the left (:117-118);
    * every frame shows the VD output of the latest sample at or before it, across gaps
      (`last_output` survives, :158-159), zeros before the first one;
    * a missing frame repeats the previous VS row once a VD output exists, otherwise both of its
      rows are zero (:170-178).
    """
    ex = np.asarray(exists, dtype=bool)
    n = ex.shape[0]
    idx = np.arange(n)
    is_sample = ex & (idx % step == 0)
    samples = idx[is_sample]
    m = samples.shape[0]
    seg_of_frame = np.cumsum(~ex)                          # segment id grows at every missing frame
    seg = seg_of_frame[samples]
    pos = np.arange(m)
    seg_start = np.zeros(m, dtype=np.int64)
    if m:
        new_seg = np.r_[True, seg[1:] != seg[:-1]]
        seg_start = np.maximum.accumulate(np.where(new_seg, pos, 0))
    windows = np.maximum(pos[:, None] - 9 + np.arange(10)[None, :], seg_start[:, None]).astype(np.int64)
    last = np.cumsum(is_sample) - 1                        # latest window index at or before frame i
    prev_existing = np.maximum.accumulate(np.where(ex, idx, -1))
    stat_src = np.where(ex, idx, np.where(last >= 0, prev_existing, -1)).astype(np.int64)
    dyn_src = last.astype(np.int64)
    return VideoPlan(samples.astype(np.int64), windows.reshape(-1, 10), stat_src, dyn_src)


@dataclass
class AudioPlan:
    starts: np.ndarray      # [Wn] window start sample
    ends: np.ndarray        # [Wn] window end sample (exclusive, clipped to L)
    f_lo: np.ndarray        # [Wn] first covered frame id
    f_hi: np.ndarray        # [Wn] one past the last covered frame id


def plan_audio(n_samples: int, fps: float, step: float = 0.5, window: int = 4, sr: int = 16000) -> AudioPlan:
    """Window schedule of get_prob_audio_8_cl.py:70-101: starts 0, step_a, ... <= L (the last window
    may be empty); frames range(round(start/sr*fps), round(end/sr*fps + 1)) with Python rounding."""
    win = window * sr
    step_a = int(step * sr)
    starts = np.arange(0, n_samples + 1, step_a, dtype=np.int64)
    ends = np.minimum(starts + win, n_samples)
    f_lo = np.rint(starts / sr * fps).astype(np.int64)
    f_hi = np.rint(ends / sr * fps + 1).astype(np.int64)
    return AudioPlan(starts, ends, f_lo, f_hi)


# =========================================================================================== engine
def balanced_batches(n: int, max_batch: int) -> List[Tuple[int, int]]:
    """[start, end) ranges of ceil(n / max_batch) batches whose sizes differ by at most one (a ragged tail batch of a few
    items costs almost a full forward's latency; at most two distinct sizes keeps the CUDA-graph cache small)."""
    if n <= 0:
        return []
    nb = -(-n // max_batch)
    base, extra = divmod(n, nb)
    out, s = [], 0
    for i in range(nb):
        e = s + base + (1 if i < extra else 0)
        out.append((s, e))
        s = e
    return out


class Engine:
    """Holds the three packed networks and runs clips through K1 -> VS -> VD, A, alignment and K4."""

    _STAGE_SLOTS = 48           # ring of pinned staging buffers: a slot is reused ~4 steps later

    def _upload(self, a: np.ndarray, dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host index array -> device through a ring of reusable pinned staging buffers, asynchronously (no stream sync,
        no cudaHostAlloc per call: a fresh pin_memory() costs more host time than the whole index plan).  A slot is only
        overwritten after the copy that last read it has completed (per-slot event; the wait is normally a no-op)."""
        a = np.ascontiguousarray(a, dtype=dtype)
        if not hasattr(self, "_stage_ring"):
            self._stage_ring = [[None, None] for _ in range(self._STAGE_SLOTS)]       # [pinned uint8 buffer, event]
            self._stage_next = 0
        slot = self._stage_ring[self._stage_next]
        self._stage_next = (self._stage_next + 1) % self._STAGE_SLOTS
        nbytes = max(a.nbytes, 1)
        if slot[0] is None or slot[0].numel() < nbytes:
            slot[0] = torch.empty(max(4096, 1 << (nbytes - 1).bit_length()), dtype=torch.uint8, pin_memory=True)
            slot[1] = torch.cuda.Event()
        else:
            slot[1].synchronize()
        host = slot[0][:a.nbytes].view(torch.from_numpy(a).dtype).view(a.shape)
        host.copy_(torch.from_numpy(a))
        if out is None:
            out = torch.empty(a.shape, dtype=host.dtype, device=self.device)
        out.copy_(host, non_blocking=True)
        slot[1].record(torch.cuda.current_stream())
        return out

    def __init__(self, sd_vs=None, sd_vd=None, sd_a=None, precision: str = "bf16", device: str = "cuda:0",
                 vs_batch: int = 256, a_batch: int = 141, use_graphs: bool = True, overlap: Optional[Tuple[int, int]] = None):
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.precision = precision
        self.vs = VSNet(sd_vs, precision, device) if sd_vs is not None else None
        self.vd = VDNet(sd_vd, precision, device) if sd_vd is not None else None
        self.a = ANet(sd_a, precision, device) if sd_a is not None else None
        self.vs_batch = vs_batch
        self.a_batch = a_batch
        self._vs_in: Optional[torch.Tensor] = None
        self._a_in: Optional[torch.Tensor] = None
        self._perm = torch.tensor(VIDEO_TO_AUDIO, device=self.device, dtype=torch.int32)
        # CUDA graphs for the two fixed-shape forwards (disabled while ops.PROFILE records per-kernel events)
        self.use_graphs = use_graphs
        # overlap = (SMs for the VS branch, SMs for the audio branch): run_clips runs the two branches side by side on
        # two streams, each persistent kernel sized to its branch's share (the VS early layers are HBM-bound, the audio
        # GEMMs tensor-bound, so the two branches complement each other); None = one after the other on one stream
        self.overlap = overlap
        vs_sms, a_sms = overlap if overlap else (0, 0)
        self._vs_sms, self._a_sms = vs_sms, a_sms
        self._a_stream = torch.cuda.Stream(device=self.device) if overlap else None
        self._vs_graph = GraphedForward(lambda x: self.vs.forward(x), vs_sms) if self.vs is not None else None
        # K1 fused into the stem (packed 224x224 crops, bf16; avcer_stem_pool_u8): the stem kernel reads the uint8 crops
        # where they lie (one eager launch, its input address changes per batch); layer1 .. fc2 replay as a graph on the
        # persistent stem output.  Bit-identical, but OFF by default: measured 240 us against 160 + 33 us for K1 + the
        # TMA-fed stem at batch 256 in three converter designs (profiles/r02_k1_fusion_negative.txt)
        self.fuse_k1 = False
        self._vs_body_graph = GraphedForward(lambda c: self.vs.body(c), vs_sms) if self.vs is not None else None
        self._vs_cat: Optional[torch.Tensor] = None
        self._a_graph = GraphedForward(lambda x: self.a.forward(x), a_sms) if self.a is not None else None
        # the VD recurrence (~45 launches) is replayed as one graph per (features buffer, window count)
        self._vd_graph = GraphedForward(lambda f, w: self.vd.forward(f, w), vs_sms) if self.vd is not None else None
        self._vd_win: Dict[int, torch.Tensor] = {}
        self._vs_out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def _vs_fwd(self, x: torch.Tensor):
        if self.use_graphs and ops.PROFILE is None:
            return self._vs_graph(x)
        with ops.sm_limit(self._vs_sms):
            return self.vs.forward(x)

    def _vd_fwd(self, feats: torch.Tensor, win_rows: np.ndarray) -> torch.Tensor:
        """win_rows: host int [10, M] rows of `feats` per time step.  The index buffer is persistent per M so that the
        captured graph's pointers stay valid; the upload happens outside the graph."""
        m = win_rows.shape[1]
        if self.use_graphs and ops.PROFILE is None:
            buf = self._vd_win.get(m)
            if buf is None:
                if len(self._vd_win) >= 16:
                    self._vd_win.pop(next(iter(self._vd_win)))
                buf = self._vd_win[m] = torch.empty((10, m), dtype=torch.int32, device=self.device)
            self._upload(win_rows, np.int32, out=buf)
            return self._vd_graph(feats, buf)
        with ops.sm_limit(self._vs_sms):
            return self.vd.forward(feats, self._upload(win_rows, np.int32))

    def _vs_outputs(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Persistent (grow-only) per-frame VS outputs: stable addresses keep the VD graph cache hot across steps."""
        if self._vs_out is None or self._vs_out[0].shape[0] < n:
            self._vs_out = (torch.empty((n, 7), device=self.device, dtype=torch.float32),
                            torch.empty((n, 512), device=self.device, dtype=torch.float32))
        return self._vs_out[0][:n], self._vs_out[1][:n]

    def _vs_fwd_u8(self, crops: torch.Tensor, consumed: Optional[torch.cuda.Event] = None):
        """crops: device uint8 [n,224,224,3] -> (probs, feats) of one batch.  `consumed` is recorded as soon as the last
        kernel that reads `crops` has been enqueued (the staging buffer may be refilled from then on)."""
        n = crops.shape[0]
        if self.fuse_k1 and self.vs.k1_fused:
            if self._vs_cat is None or self._vs_cat.shape[0] < n:
                self._vs_cat = torch.empty((max(n, self.vs_batch), 55, 55, 128), device=self.device, dtype=self.vs.dtype)
            cat = self._vs_cat[:n]
            with ops.sm_limit(self._vs_sms):
                self.vs.stem_u8(crops, cat)
                if consumed is not None:
                    consumed.record(torch.cuda.current_stream())
                if self.use_graphs and ops.PROFILE is None:
                    return self._vs_body_graph(cat)
                return self.vs.body(cat)
        x = self._vs_input(n)
        ops.preprocess(crops, n, x, self.vs.input_layout)
        if consumed is not None:
            consumed.record(torch.cuda.current_stream())
        return self._vs_fwd(x)

    def _a_fwd(self, x: torch.Tensor):
        if self.use_graphs and ops.PROFILE is None:
            return self._a_graph(x)
        with ops.sm_limit(self._a_sms):
            return self.a.forward(x)

    # ------------------------------------------------------------------ VS over packed 224x224 crops
    def _vs_input(self, n: int) -> torch.Tensor:
        if self._vs_in is None or self._vs_in.shape[0] < n:
            self._vs_in = self.vs.alloc_input(max(n, self.vs_batch))
        return self._vs_in[:n]

    def vs_forward_u8(self, crops_u8: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """crops_u8: device uint8 [n,224,224,3] BGR.  Returns (probs [n,7] fp32, features [n,512])."""
        n = crops_u8.shape[0]
        probs, feats = self._vs_outputs(n)
        for s, e in balanced_batches(n, self.vs_batch):
            p, f = self._vs_fwd_u8(crops_u8[s:e])
            probs[s:e].copy_(p)
            feats[s:e].copy_(f)
        return probs, feats

    def vs_forward_host(self, crops_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same as vs_forward_u8 for crops in pinned HOST memory: the H2D copy of batch i+1 runs on a
        copy stream while batch i is preprocessed and classified (double-buffered device staging)."""
        n = crops_host.shape[0]
        bs = self.vs_batch
        probs, feats = self._vs_outputs(n)
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty((bs, 224, 224, 3), device=self.device, dtype=torch.uint8) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream()
        ranges = balanced_batches(n, bs)
        starts = [r[0] for r in ranges]

        def issue_copy(i):
            s0, e0 = ranges[i]
            k = i & 1
            if i >= 2:
                self._copy_stream.wait_event(self._free[k])
            else:
                self._copy_stream.wait_stream(cur)
            with torch.cuda.stream(self._copy_stream):
                self._stage[k][: e0 - s0].copy_(crops_host[s0:e0], non_blocking=True)
                self._ready[k].record(self._copy_stream)

        if starts:
            issue_copy(0)
        for i, (s0, e0) in enumerate(ranges):
            if i + 1 < len(starts):
                issue_copy(i + 1)
            k = i & 1
            cur.wait_event(self._ready[k])
            p, f = self._vs_fwd_u8(self._stage[k][: e0 - s0], consumed=self._free[k])
            probs[s0:e0].copy_(p)
            feats[s0:e0].copy_(f)
        return probs, feats

    def vs_forward_ragged(self, flat_u8: torch.Tensor, offsets: np.ndarray, heights: np.ndarray, widths: np.ndarray,
                          out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """Crops of arbitrary size packed back to back in `flat_u8` (K1 does the NEAREST resize).  `out`: (probs [n,7],
        feats [n,512]) slices to fill instead of the engine's own result buffers (a clip processed in chunks)."""
        n = len(offsets)
        probs, feats = out if out is not None else self._vs_outputs(n)
        off = self._upload(np.asarray(offsets), np.int64)
        hh = self._upload(np.asarray(heights), np.int32)
        ww = self._upload(np.asarray(widths), np.int32)
        for s in range(0, n, self.vs_batch):
            e = min(n, s + self.vs_batch)
            x = self._vs_input(e - s)
            ops.preprocess(flat_u8, e - s, x, self.vs.input_layout, offsets=off[s:e], heights=hh[s:e], widths=ww[s:e])
            p, f = self._vs_fwd(x)
            probs[s:e].copy_(p)
            feats[s:e].copy_(f)
        return probs, feats

    # ------------------------------------------------------------------ video branch of a set of clips
    def video_rows(self, probs: torch.Tensor, feats: torch.Tensor, exists_list: Sequence[np.ndarray],
                   fps_list: Sequence[float], stat_out: Optional[torch.Tensor] = None,
                   dyn_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, List[VideoPlan]]:
        """probs/feats hold the VS outputs of the present frames of all clips, clip after clip.
        Returns per-frame (stat [sumN,7] VS probabilities, dyn [sumN,7] VD logits) in VIDEO_ORDER
        with the reference's carry-forward / gap semantics, plus the per-clip plans."""
        plans = []
        stat_idx, dyn_idx, win_all, sample_rows = [], [], [], []
        present_base = 0      # row offset of this clip in probs/feats
        sample_base = 0       # offset of this clip's samples in the gathered unique-feature list
        for ex, fps in zip(exists_list, fps_list):
            ex = np.asarray(ex, dtype=bool)
            plan = plan_video(ex, vd_step(fps))
            plans.append(plan)
            row_of_frame = np.cumsum(ex) - 1                           # present-row index of an existing frame
            stat_idx.append(np.where(plan.stat_src >= 0, present_base + row_of_frame[np.maximum(plan.stat_src, 0)], -1))
            dyn_idx.append(np.where(plan.dyn_src >= 0, sample_base + plan.dyn_src, -1))
            win_all.append(plan.windows + sample_base)
            sample_rows.append(present_base + row_of_frame[plan.samples])
            present_base += int(ex.sum())
            sample_base += plan.samples.shape[0]
        dev = self.device
        stat_idx_t = self._upload(np.concatenate(stat_idx), np.int32)
        dyn_idx_t = self._upload(np.concatenate(dyn_idx), np.int32)
        n_total = stat_idx_t.numel()
        stat = ops.gather_rows(probs, stat_idx_t, n_total, out=stat_out)
        windows = np.concatenate(win_all, axis=0) if win_all else np.zeros((0, 10), dtype=np.int64)
        if windows.shape[0]:
            # windows index the sample list; map them to rows of `feats` so no feature copy is needed
            rows = np.concatenate(sample_rows)
            vd_logits = self._vd_fwd(feats, rows[windows].T)       # [10, M] feature rows per time step
        else:
            vd_logits = torch.zeros((1, 7), device=dev, dtype=torch.float32)
        dyn = ops.gather_rows(vd_logits, dyn_idx_t, n_total, out=dyn_out)
        return stat, dyn, plans

    # ------------------------------------------------------------------ audio branch
    def audio_window_logits(self, wav: torch.Tensor, starts: np.ndarray, ends: np.ndarray, padding: str, win: int) -> torch.Tensor:
        """wav: device fp32 buffer (one clip, or several clips back to back with global starts/ends).
        Returns per-window logits [Wn, ncls] fp32."""
        if padding == "repeat" and bool((ends - starts == 0).any()):
            raise ZeroDivisionError("integer division or modulo by zero")      # data/utils.py:66 on the empty tail window
        if padding not in ops.PAD_MODES:
            raise UnboundLocalError("cannot access local variable 'a_fss' where it is not associated with a value")
        st = self._upload(starts, np.int64)
        en = self._upload(ends, np.int64)
        wn = int(st.numel())
        out = torch.empty((wn, self.a.num_classes), device=self.device, dtype=torch.float32)
        for s, e in balanced_batches(wn, self.a_batch):
            if self._a_in is None or self._a_in.shape[0] < self.a_batch or self._a_in.shape[1] != win:
                self._a_in = torch.empty((self.a_batch, win), device=self.device, dtype=torch.float32)
            x = ops.audio_normalize_windows(wav, st[s:e], win, padding, ends=en[s:e], out=self._a_in[:e - s])
            out[s:e].copy_(self._a_fwd(x))
        return out

    def audio_frame_means(self, logits: torch.Tensor, f_lo: np.ndarray, f_hi: np.ndarray, n_frames: int) -> torch.Tensor:
        lo = self._upload(f_lo, np.int32)
        hi = self._upload(f_hi, np.int32)
        return ops.window_to_frame_mean(logits, lo, hi, n_frames)

    def audio_rows(self, wav_cat: torch.Tensor, wav_lens: Sequence[int], fps_list: Sequence[float], n_frames: Sequence[int],
                   step: float = 0.5, window: int = 4, sr: int = 16000, padding: str = "mean", out: Optional[torch.Tensor] = None):
        """Audio branch of several clips whose waveforms are concatenated in `wav_cat`.  Returns
        (per-frame mean logits [sum N, ncls] with the tail rule of run.py:99-103 applied, window logits)."""
        base = np.r_[0, np.cumsum(n_frames)]
        woff = np.r_[0, np.cumsum(wav_lens)]
        st_all, en_all, lo_all, hi_all, tail_src = [], [], [], [], []
        for ci, (L, fps) in enumerate(zip(wav_lens, fps_list)):
            ap = plan_audio(int(L), fps, step, window, sr)
            st_all.append(ap.starts + woff[ci])
            en_all.append(ap.ends + woff[ci])
            nf = n_frames[ci]
            lo_all.append(base[ci] + np.minimum(ap.f_lo, nf))
            hi_all.append(base[ci] + np.minimum(ap.f_hi, nf))          # frame ids >= N are dropped by the isin filter (run.py:96)
            covered = int(min(nf, ap.f_hi.max()))
            src = np.arange(nf)
            src[covered:] = max(covered - 1, 0)
            tail_src.append(base[ci] + src)
        logits = self.audio_window_logits(wav_cat, np.concatenate(st_all), np.concatenate(en_all), padding, window * sr)
        a_mean = self.audio_frame_means(logits, np.concatenate(lo_all), np.concatenate(hi_all), int(base[-1]))
        tail = self._upload(np.concatenate(tail_src), np.int32)
        return ops.gather_rows(a_mean, tail, int(base[-1]), out=out), logits

    # ------------------------------------------------------------------ K4 on aligned per-frame rows
    def fuse(self, stat_video_order: torch.Tensor, dyn_video_order: torch.Tensor, audio_mean_logits: torch.Tensor,
             weights_1, weights_2, ce_weights_type: bool, ce_mask: bool, f64_video: bool = False,
             labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """run.py:85-165 on device: permute the video columns into audio order, softmax the VD logits and
        the audio mean logits (first 7 classes), then K4.  Returns int64 labels [4, n].
        f64_video: the reference's video tables are float64 for this clip (a zero row was appended, np.array promotion at
        get_prob_video.py:89,182-187), so the VD softmax and the fusion run in float64; the audio table stays float32."""
        n = stat_video_order.shape[0]
        p_vs = ops.gather_rows(stat_video_order, None, n, perm=self._perm)
        dyn = ops.gather_rows(dyn_video_order, None, n, perm=self._perm)
        p_a = ops.softmax7(audio_mean_logits)
        if f64_video:
            p_vs, p_vd, p_a = p_vs.double(), ops.softmax7(dyn.double()), p_a.double()
        else:
            p_vd = ops.softmax7(dyn)
        return ops.fuse_compound(p_vs, p_vd, p_a, weights_1, weights_2, ce_weights_type, ce_mask, labels=labels)

    # ------------------------------------------------------------------ whole clips, batched
    def fuse_clips(self, stat: torch.Tensor, dyn: torch.Tensor, a_rows: torch.Tensor, f64_flags: Sequence[bool],
                   n_frames: Sequence[int], weights_1, weights_2, ce_weights_type: bool, ce_mask: bool,
                   labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K4 over the per-frame rows of several clips (clip after clip).  Clips whose video tables the reference holds in
        float64 (`f64_flags`: a zero row was appended -- frames before the first VD output or without any crop) get their
        VD softmax and fusion redone in float64, like numpy does for them.  `labels`: optional [4, n] destination, which
        may be a column slice of a wider buffer."""
        labels = self.fuse(stat, dyn, a_rows, weights_1, weights_2, ce_weights_type, ce_mask, labels=labels)
        base = 0
        for flag, nf in zip(f64_flags, n_frames):
            if nf and flag:
                sl = slice(base, base + nf)
                self.fuse(stat[sl], dyn[sl], a_rows[sl], weights_1, weights_2, ce_weights_type, ce_mask, f64_video=True,
                          labels=labels[:, sl])
            base += nf
        return labels

    @staticmethod
    def needs_f64(exists: np.ndarray, fps: float) -> bool:
        """True when the reference's video DataFrames of this clip are float64: a zero row is appended for every frame
        before the first VD output (get_prob_video.py:89,163-178), i.e. unless frame 0 has a crop."""
        ex = np.asarray(exists, dtype=bool)
        return bool(ex.size) and not bool(ex[0])

    def run_clips(self, crops_u8: torch.Tensor, exists_list: Sequence[np.ndarray], fps_list: Sequence[float],
                  wav_cat: torch.Tensor, wav_lens: Sequence[int], weights_1, weights_2, ce_weights_type: bool, ce_mask: bool,
                  step: float = 0.5, window: int = 4, sr: int = 16000, padding: str = "mean",
                  rows_out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                  fuse: bool = True) -> Dict[str, torch.Tensor]:
        """All clips at once.  crops_u8: uint8 [sum present frames, 224,224,3] BGR; wav_cat: fp32 waveforms
        back to back.  Host (pinned) tensors are copied to the device first; device tensors are used as is.
        rows_out = (stat [n,7], dyn [n,7], audio_mean [n,ncls]): contiguous destinations of the per-frame rows (slices of
        an all-gather send buffer, dist.ShardedRunner); fuse=False stops before K4 (the caller fuses the gathered rows)."""
        if not wav_cat.is_cuda:
            wav_cat = wav_cat.to(self.device, non_blocking=True)
        n_frames = [len(e) for e in exists_list]
        so, do, ao = rows_out if rows_out is not None else (None, None, None)
        if self.overlap and ops.PROFILE is None:
            cur = torch.cuda.current_stream()
            self._a_stream.wait_stream(cur)
            with torch.cuda.stream(self._a_stream):                      # audio branch: enqueued first, runs beside VS / VD
                a_rows, logits = self.audio_rows(wav_cat, wav_lens, fps_list, n_frames, step, window, sr, padding, out=ao)
            probs, feats = self.vs_forward_u8(crops_u8) if crops_u8.is_cuda else self.vs_forward_host(crops_u8)
            stat, dyn, plans = self.video_rows(probs, feats, exists_list, fps_list, so, do)
            cur.wait_stream(self._a_stream)
            for t in (a_rows, logits, wav_cat):
                t.record_stream(cur)
            wav_cat.record_stream(self._a_stream)
        else:
            probs, feats = self.vs_forward_u8(crops_u8) if crops_u8.is_cuda else self.vs_forward_host(crops_u8)
            stat, dyn, plans = self.video_rows(probs, feats, exists_list, fps_list, so, do)
            a_rows, logits = self.audio_rows(wav_cat, wav_lens, fps_list, n_frames, step, window, sr, padding, out=ao)
        out = {"stat": stat, "dyn": dyn, "audio_mean": a_rows, "window_logits": logits}
        if fuse:
            flags = [bool((p.stat_src < 0).any()) or bool((p.dyn_src < 0).any()) for p in plans]
            out["labels"] = self.fuse_clips(stat, dyn, a_rows, flags, n_frames, weights_1, weights_2, ce_weights_type, ce_mask)
        return out
