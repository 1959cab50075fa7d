#!/usr/bin/env python
"""Benchmark of the batched inference-and-fusion path (BASELINE.json metric: end-to-end frames/sec,
VS+VD+A+fusion).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A *step* = one pass of the whole path over the rank's shard of synthetic clips (BASELINE config 4,
weak scaling: `--clips-per-gpu` clips of 60 s / 25 fps = 1500 face crops 224x224 + 960 000 audio
samples each).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the public
call (Engine.run_clips) with pinned HOST buffers, H2D of crops + waveforms and D2H of the labels inside
the timed region.  The inputs of one step (>= 1.8 GB of crops) are far larger than the 126 MB L2, so
successive steps cannot be served from cache.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME_VS = 7.667e9
METRIC = "end_to_end_frames_per_sec_vs_vd_a_fusion"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region (pynvml, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
# One CPU "step" = a bounded sample of the bench workload with the workload's own proportions: a config-4 clip has
# 1500 frames, 300 VD windows and 121 audio windows, i.e. 12.4 frames per audio window; the sample keeps that ratio, so
# frames / measured wall seconds of the sample is the same frames/s metric as the GPU arm's.  Nothing is extrapolated:
# `ms_per_step` is the measured wall time of the step that was really executed.
CPU_SAMPLE_FRAMES = 75
CPU_SAMPLE_WINDOWS = 6


class CpuSample:
    """Inputs + weights of the CPU arms, built once (outside the timed steps)."""

    def __init__(self, frames: int = CPU_SAMPLE_FRAMES, windows: int = CPU_SAMPLE_WINDOWS):
        from avcer_b200 import get_weights_matrices as gwm, synthetic as syn

        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.frames, self.windows = frames, windows
        self.sd_vs, self.sd_vd = syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1)
        self.sd_a = syn.make_audio_state_dict(2, 8, "spread", 12)
        self.crops = syn.make_crops(1, frames)
        # `windows` windows of 4 s at step 0.5 s: L just below a multiple of the step (no empty tail window)
        self.wav = syn.make_wav(2, int(16000 * 0.5 * windows) - 160)
        self.w1, self.w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]

    def batched_step(self) -> float:
        """The oracle port of the reference algorithm (torch CPU fp32, all host threads), batched the way a CPU user would
        batch it (32 crops / 4 windows per forward): VS + VD, A, alignment, fusion of one sample.  Returns wall seconds."""
        import pandas as pd

        from oracle import audio as oa, fusion as of, video as ov

        t0 = time.perf_counter()
        dyn, stat = ov.predict_video(list(self.crops), 25, self.sd_vs, self.sd_vd)
        rows, ids, logits = oa.predict_audio(self.wav, 25, self.sd_a)
        assert logits.shape[0] == self.windows
        stat_df, dyn_df = pd.DataFrame(stat, columns=of.VIDEO_ORDER), pd.DataFrame(dyn, columns=of.VIDEO_ORDER)
        audio_df = pd.DataFrame(rows, columns=of.AUDIO_ORDER)
        audio_df["frames"] = [str(i).zfill(6) + ".jpg" for i in ids]
        of.get_c_expr_db_pred(stat_df, dyn_df, audio_df, "c", self.w1, self.w2, False, True)
        return time.perf_counter() - t0

    def loop_step(self, frames: int = 25, windows: int = 2) -> dict:
        """The reference's own execution shape (oracle/loop.py: JPEG decode + PIL preprocessing + batch-1 forwards per frame,
        batch-1 forward per audio window) on a smaller sample with the same frames-per-window ratio."""
        import tempfile

        import cv2

        from oracle import loop as ol

        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "clip", "00"))
            for i in range(frames):
                cv2.imwrite(os.path.join(td, "clip", "00", f"{i:06d}.jpg"), self.crops[i % len(self.crops)])
            t0 = time.perf_counter()
            ol.video_loop(os.path.join(td, "clip"), 25, frames, self.sd_vs, self.sd_vd)
            t_v = time.perf_counter() - t0
        wav = self.wav[: int(16000 * 0.5 * windows) - 160]
        t0 = time.perf_counter()
        ol.audio_loop(wav, 25, self.sd_a)
        t_a = time.perf_counter() - t0
        return {"value": frames / (t_v + t_a), "unit": "frames/s", "sample": f"{frames} JPEG crops (imread + PIL + batch-1 VS, VD on every 5th) "
                f"+ {windows} audio windows at batch 1, measured wall {t_v + t_a:.2f} s", "ms_per_frame_video": 1e3 * t_v / frames,
                "ms_per_window_audio": 1e3 * t_a / windows}

    def describe(self) -> str:
        return (f"oracle port of the reference algorithm (torch CPU fp32, batched 32 crops / 4 windows): one step = VS+VD on "
                f"{self.frames} crops + A on {self.windows} windows of 4 s + alignment + fusion (the 12.4 frames per audio window of a "
                f"config-4 clip), wall-clocked as run; nothing scaled")


def cpu_baseline(reps: int = 3):
    """~10-30 s of CPU work on the box's host cores: `reps` measured steps of the batched port (median) and one pass of the
    batch-1 loop (the reference's own execution shape)."""
    smp = CpuSample()
    smp.batched_step()                                                             # warm-up (thread pools, allocator)
    walls = [smp.batched_step() for _ in range(reps)]
    w = statistics.median(walls)
    return {"value": smp.frames / w, "unit": "frames/s", "cores": smp.cores, "kind": "port", "sample": smp.describe(),
            "ms_per_step": 1e3 * w, "steps_measured": reps, "batch1_loop": smp.loop_step()}


def run_reference_arm(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    smp = CpuSample()
    walls = []
    for i in range(args.warmup + args.steps):
        w = smp.batched_step()
        if i >= args.warmup:
            walls.append(w)
    w = statistics.median(walls)
    v = smp.frames / w
    base = {"value": v, "unit": "frames/s", "cores": smp.cores, "kind": "port", "sample": smp.describe(),
            "ms_per_step": 1e3 * w, "wall_s_timed_steps": sum(walls), "batch1_loop": smp.loop_step()}
    cfg = dict(workload, cpu_step=f"{smp.frames} frames + {smp.windows} audio windows per step (bounded sample of the workload)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * w, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "one host (all its cores) regardless of --gpus: compare with the N=1 line only"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--clips", type=int, default=64)              # strong scaling: the fixed clip set
    ap.add_argument("--clips-per-gpu", type=int, default=8)       # weak scaling
    ap.add_argument("--clip-seconds", type=int, default=60)
    ap.add_argument("--fps", type=int, default=25)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    # windows per audio forward: 141 x 199 tokens = 110 row tiles of 256, i.e. 440 / 1320 / 1760 tiles for the encoder's N = 1024 /
    # 3072 / 4096 GEMMs = 5.95 / 17.8 / 23.8 waves over the 74 CTA pairs (64 windows: 2.70 / 8.1 / 10.8 -> a tenth of every
    # GEMM is a partly empty wave); measured 51.0 k -> 51.7 k frames/s
    ap.add_argument("--a-batch", type=int, default=141)
    ap.add_argument("--vs-batch", type=int, default=1536)     # crops per VS forward inside the pipeline (config 2 below stays at 256); 768 - 2048 measured within 1 %
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()

    from avcer_b200.pipeline import plan_audio

    # ---- the clip set.  strong (default, BASELINE config 4: a FIXED set of clips sharded over the GPUs): --clips clips of
    # mixed length, mean --clip-seconds; weak: --clips-per-gpu clips of --clip-seconds on every rank (round-1 workload).
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scaling == "strong":
        pattern = (1.0, 0.5, 1.5, 1.0, 2.0, 1.0 / 6.0, 1.0, 5.0 / 6.0)                  # x clip_seconds; mean 1.0
        durations = [max(1, int(round(args.clip_seconds * pattern[i % len(pattern)]))) for i in range(args.clips)]
    else:
        durations = [args.clip_seconds] * (args.clips_per_gpu * world_env)
    frames_all = [d * args.fps for d in durations]
    samples_all = [d * 16000 for d in durations]
    windows_all = [len(plan_audio(sm, args.fps).starts) for sm in samples_all]
    total_frames = sum(frames_all)
    n_frames, n_windows = args.clip_seconds * args.fps, len(plan_audio(args.clip_seconds * 16000, args.fps).starts)
    what = (f"BASELINE config 4, fixed set sharded over the GPUs (strong scaling): {len(durations)} clips of "
            f"{min(durations)}-{max(durations)} s (mean {sum(durations) / len(durations):.0f} s) @ {args.fps} fps = {total_frames} face crops 224x224 + "
            f"{sum(durations)} s of 16 kHz audio ({sum(windows_all)} windows of 4 s / step 0.5 s) per step"
            if args.scaling == "strong" else
            f"BASELINE config 4 shard (weak scaling): {args.clips_per_gpu} clips/GPU x {args.clip_seconds} s @ {args.fps} fps "
            f"({n_frames} crops 224x224 + {n_windows} windows of 4 s / step 0.5 s per clip)")
    workload = {"workload": what + ", VS ResNet-50 + VD LSTM + A wav2vec2-L12 (8 classes) + fusion Rule 1 with the AV-8cl weight matrix",
                "clips": len(durations), "frames_per_step": total_frames, "windows_per_step": sum(windows_all),
                "frames_per_clip": n_frames, "windows_per_clip": n_windows,
                "vs_batch": args.vs_batch, "a_batch": args.a_batch, "parallelism": f"clip-sharded x{args.gpus} (LPT on frames + 11.6 x windows)",
                "l2_policy": "inputs (>= 1.8 GB per rank and step) larger than L2", "weights": "random-init, seeded"}
    if args.impl == "reference":
        return run_reference_arm(args, workload)

    from avcer_b200 import dist as adist, get_weights_matrices as gwm, ops, synthetic as syn
    from avcer_b200.pipeline import Engine

    rank, world, local_rank = adist.init_from_env()
    dev = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    peaks = load_peaks()
    eng = Engine(syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12),
                 precision=args.precision, device=dev, vs_batch=args.vs_batch, a_batch=args.a_batch)
    w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]

    # every rank knows the metadata of all clips and generates the data of its own shard (seeded by the global clip index)
    runner = adist.ShardedRunner(eng, frames_all, windows_all, [False] * len(durations), rank, world)
    mine = runner.my_clips
    local_frames = sum(frames_all[i] for i in mine)
    crops_dev = torch.empty((local_frames, 224, 224, 3), dtype=torch.uint8, device=dev)
    wav_dev = torch.empty(sum(samples_all[i] for i in mine), dtype=torch.float32, device=dev)
    fo = so = 0
    for i in mine:
        g = torch.Generator(device=dev).manual_seed(1000 + i)
        crops_dev[fo: fo + frames_all[i]] = torch.randint(0, 256, (frames_all[i], 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
        wav_dev[so: so + samples_all[i]] = torch.randn(samples_all[i], device=dev, generator=g) * 0.1
        fo += frames_all[i]
        so += samples_all[i]
    exists = [np.ones(frames_all[i], dtype=bool) for i in mine]
    fps_list = [float(args.fps)] * len(mine)
    wav_lens = [samples_all[i] for i in mine]
    c = len(mine)

    def step(crops, wav):
        # K1 -> VS -> VD, A, alignment on this rank's clips; per-frame rows straight into the all-gather send buffer; ONE
        # collective; the fusion tail (permute, softmax, K4) over every gathered block -> labels of all frames on all ranks
        return runner.step(crops, exists, fps_list, wav, wav_lens, w1, w2, False, True)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(crops_dev, wav_dev)
    barrier()
    # Host cost of enqueuing ONE step into an empty launch queue (the timed loop below enqueues back to back, where the
    # host blocks on the driver's bounded queue and its wall time only mirrors the GPU's)
    t_iso = time.perf_counter()
    step(crops_dev, wav_dev)
    host_isolated_ms = (time.perf_counter() - t_iso) * 1e3
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.STATS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = []
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step(crops_dev, wav_dev)
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    barrier()
    launches = ops.STATS["launches"] - launches0
    ms = e0.elapsed_time(e1)
    # Per-kernel CUDA-event timing needs individually launched kernels; the timed steps replay CUDA graphs
    # (one graph per 256-crop VS batch / 64-window A batch).  One more step of the same workload is run
    # eagerly right after the timed region, on the same stream, with an event pair around every launch.
    ops.PROFILE = prof
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step(crops_dev, wav_dev)
    p1.record()
    ops.PROFILE = None
    barrier()
    profiled_step_ms = p0.elapsed_time(p1)
    clocks = sampler.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = total_frames * args.steps / (ms / 1e3)

    # per-kernel roofline from the profiled step: every launch carries (kernel, algorithmic work, algorithmic bytes)
    agg = {}
    for tag, work, a, b, nbytes in prof:
        d = agg.setdefault(tag, [0.0, 0.0, 0, 0.0])
        d[0] += work
        d[1] += a.elapsed_time(b)
        d[2] += 1
        d[3] += nbytes
    kern = {k: {"work": v[0], "ms": v[1], "launches": v[2], "bytes": v[3]} for k, v in agg.items()}
    step_ms = ms / args.steps
    roofline = None
    prof_note = ("per-launch CUDA events on one extra eager step run right after the timed region "
                 f"(that step: {profiled_step_ms:.1f} ms; timed steps replay CUDA graphs: {step_ms:.1f} ms)")
    # NCU names of the contraction kernels (profiles/r02_kernel_digest.json: dram__bytes_read.sum + dram__bytes_write.sum per launch)
    ncu_names = {"tc_gemm2_kernel<OUT_TMA,256,ring>": "tc_gemm2_kernel<0, 256, 4, 0>", "tc_gemm2_kernel<OUT_TMA_RES,256,ring>": "tc_gemm2_kernel<1, 256, 4, 0>",
                 "tc_gemm2_kernel<OUT_TMA,256,FLAT>": "tc_gemm2_kernel<0, 256, 4, 1>", "tc_gemm2_kernel<OUT_TMA_RES,256,FLAT>": "tc_gemm2_kernel<1, 256, 4, 1>",
                 "conv3x3_kernel<64>": "conv3x3_kernel<64, 1>", "conv3x3_kernel<128>": "conv3x3_kernel<128, 0>"}
    digest_path = os.path.join(ROOT, "profiles", "r02_kernel_digest.json")
    digest = json.load(open(digest_path)) if os.path.exists(digest_path) else {}
    fam = {k: v for k, v in kern.items() if k.startswith("contract")}
    if fam:
        name, tc = max(fam.items(), key=lambda kv: kv[1]["ms"])          # the ONE dominant kernel of the step
        ach = tc["work"] / (tc["ms"] / 1e3) / 1e12
        fam_work, fam_ms, fam_n = sum(v["work"] for v in fam.values()), sum(v["ms"] for v in fam.values()), sum(v["launches"] for v in fam.values())
        short = name.split(":", 1)[-1]
        dg = digest.get(ncu_names.get(short, short), {})
        roofline = {"kernel": short + " (two-SM tcgen05 implicit GEMM, 256x256 tiles: FFN / qkv / conv GEMMs of the audio network, wide VS layers)"
                    if short.startswith("tc_gemm2_kernel<OUT_TMA,256,ring") else short,
                    "bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                    "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
                    "traffic": dg.get("dram_bytes_per_launch"),
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel (ncu --set full, profiles/r02_kernel_digest.json); "
                                    "algorithmic_bytes_per_launch = operands + output once" if dg else "no ncu capture of this kernel in profiles/",
                    "algorithmic_bytes_per_launch": tc["bytes"] / tc["launches"], "algorithmic_flop_per_launch": tc["work"] / tc["launches"],
                    "launches_per_step": tc["launches"], "avg_launch_us": 1e3 * tc["ms"] / tc["launches"],
                    "share_of_step": tc["ms"] / profiled_step_ms, "how": prof_note,
                    "all_contraction_kernels": {"achieved": fam_work / (fam_ms / 1e3) / 1e12, "frac": fam_work / (fam_ms / 1e3) / 1e12 / peaks["tf_sustained"],
                                                "launches_per_step": fam_n, "share_of_step": fam_ms / profiled_step_ms,
                                                "per_kernel": {k.split(":", 1)[-1]: {"tflops": v["work"] / (v["ms"] / 1e3) / 1e12, "ms": v["ms"], "launches": v["launches"]}
                                                               for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}}
    extra = {}
    for tag, name in (("preprocess", "k1_preprocess"), ("fuse_compound", "k4_fusion")):
        if tag in kern:
            gbs = kern[tag]["work"] / (kern[tag]["ms"] / 1e3) / 1e9
            extra[name] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                           "launches_per_step": kern[tag]["launches"], "avg_launch_us": 1e3 * kern[tag]["ms"] / kern[tag]["launches"]}

    # end to end through the public call with pinned host buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.skip_e2e:
        crops_host = torch.empty(crops_dev.shape, dtype=torch.uint8, pin_memory=True)
        crops_host.copy_(crops_dev)
        wav_host = torch.empty(wav_dev.shape, dtype=torch.float32, pin_memory=True)
        wav_host.copy_(wav_dev)
        del crops_dev
        torch.cuda.empty_cache()
        labels_host = None
        for _ in range(max(1, args.warmup - 1)):
            labels_host = step(crops_host, wav_host).cpu()
        barrier()
        e0.record()
        for _ in range(args.steps):
            labels_host = step(crops_host, wav_host).cpu()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e = {"value": total_frames * args.steps / (float(t.item()) / 1e3), "unit": "frames/s",
               "h2d_bytes_per_step": int(total_frames * 224 * 224 * 3 + sum(samples_all) * 4),      # whole job: every rank's crops + waveforms
               "d2h_bytes_per_step": int(labels_host.numel() * 8 * world),                        # every rank reads all labels back
               "api": "avcer_b200.dist.ShardedRunner.step -> pipeline.Engine.run_clips (pinned host inputs)"}
        crops_dev = crops_host.to(dev)

    # BASELINE config 2: VS alone, batch 256, bf16 (K1 + ResNet-50), L2 flushed between iterations
    vs_alone = None
    if rank == 0:
        x256 = crops_dev[:256].contiguous()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        times = []
        for i in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.vs_forward_u8(x256)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(a.elapsed_time(b))
        tm = statistics.median(times)
        tf = 256 * FLOP_PER_FRAME_VS / (tm / 1e3) / 1e12
        vs_alone = {"workload": "VS ResNet-50 alone, batch 256 crops 224x224, K1 + forward", "ms": tm, "frames_per_s": 256 / (tm / 1e3),
                    "tflops": tf, "frac_of_bf16_burst_peak": tf / peaks["tf_burst"], "l2_policy": "256 MB flush between iterations"}

    # K1 / K4 at their full-size operating points (the step above launches them on 256-crop batches and on one
    # shard's 12 000 frames, where K4 is launch-bound): 1024 crops for K1, BASELINE config 4's 1.5 M frames for K4.
    micro = {}
    if rank == 0:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush2 = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
        xin = eng.vs.alloc_input(1024)
        src1k = crops_dev[:1024].contiguous()
        nfr = 1_500_000
        gk = torch.Generator(device=dev).manual_seed(7)
        ps = [torch.softmax(torch.randn(nfr, 7, device=dev, generator=gk), 1).contiguous() for _ in range(3)]
        lab = torch.empty((4, nfr), device=dev, dtype=torch.int64)

        def timed(fn, reps=8):
            """Short kernels: the Python/ctypes call costs more host time than the kernel runs, so CUDA events around
            a single eager launch would time the host.  `reps` back-to-back launches are captured in a CUDA graph and
            the replay is timed; the working set (K1: 0.58 GB, K4: 0.17 GB) is larger than L2."""
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            gc.collect()                   # no finaliser (e.g. of an old CUDA graph) may run while the stream is capturing
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    for _ in range(reps):
                        fn()
            finally:
                gc.enable()
            ts = []
            for i in range(5):
                flush.zero_()
                flush2.sum()           # L2 holds unrelated clean lines when the replay starts
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(b) / reps)
            return statistics.median(ts)

        t1 = timed(lambda: ops.preprocess(src1k, 1024, xin, eng.vs.input_layout))
        t4 = timed(lambda: ops.fuse_compound(ps[0], ps[1], ps[2], w1, w2, False, True, labels=lab))
        for name, t, work, unit_desc in (("k1_preprocess_1024_crops", t1, 1024 * 451584.0, "451584 B/frame"),
                                         ("k4_fusion_1p5M_frames", t4, nfr * 116.0, "116 B/frame")):
            gbs = work / (t / 1e3) / 1e9
            micro[name] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                           "ms": t, "algorithmic": unit_desc, "l2_policy": "working set larger than L2; L2 flushed before the 8-launch graph replay",
                           "peak_source": f"{peaks['source']} HBM copy"}
        del ps, lab, xin, flush, flush2
        # baseline-JPEG decode of the crops on the GPU (the reference: cv2.imread per frame on one core)
        try:
            import cv2

            from avcer_b200 import jpeg as ajpeg

            base = syn.make_crops(3, 30)
            files = [cv2.imencode(".jpg", base[i % 30])[1].tobytes() for i in range(1500)]
            ajpeg.decode_batch(files, dev)
            torch.cuda.synchronize()
            walls, devs = [], []
            for _ in range(3):
                ajpeg.PROFILE = []
                t0 = time.perf_counter()
                ajpeg.decode_batch(files, dev)
                torch.cuda.synchronize()
                walls.append((time.perf_counter() - t0) * 1e3)
                devs.append(sum(a.elapsed_time(b) for a, b in ajpeg.PROFILE))
            ajpeg.PROFILE = None
            t0 = time.perf_counter()
            for f in files[:100]:
                cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
            t_cv = (time.perf_counter() - t0) / 100
            micro["jpeg_decode_1500_crops"] = {"bound": "latency (one thread per image in the Huffman kernel)", "crops_per_s_wall": 1500 / min(walls) * 1e3,
                                               "crops_per_s_kernels": 1500 / min(devs) * 1e3, "ms_wall": min(walls), "ms_kernels": min(devs),
                                               "bytes_per_file": sum(map(len, files)) / 1500, "cv2_imdecode_crops_per_s_one_core": 1 / t_cv,
                                               "parity": "bit-identical to cv2.imread (tests/test_gpu_preprocess.py)"}
        except Exception as e:  # pragma: no cover  (cv2 missing: the figure is optional)
            micro["jpeg_decode_1500_crops"] = {"skipped": f"{type(e).__name__}: {e}"}

        # the file-based drop-in call a user of the reference makes for one config-4 clip: 1500 JPEG crops on disk ->
        # get_prob_video.preprocess_video_and_predict (listdir, file reads, GPU JPEG decode, K1, VS, VD, DataFrames)
        try:
            import tempfile

            import cv2

            from avcer_b200 import config as acfg, get_prob_video as gpv

            with tempfile.TemporaryDirectory(prefix="avcer_bench_") as td:
                os.makedirs(os.path.join(td, "clip", "00"))
                # face crops as a detector writes them: every file its own size (here 160 .. 319 px, cut out of 50 synthetic
                # faces), so K1 resizes and the JPEG headers differ from file to file
                base = syn.make_crops(5, 50, 320)
                rs = np.random.default_rng(11)
                for i in range(1500):
                    hh, ww = (int(v) for v in rs.integers(160, 320, 2))
                    cv2.imwrite(os.path.join(td, "clip", "00", f"{i:06d}.jpg"), np.ascontiguousarray(base[i % 50][:hh, :ww]))
                acfg.set_precision(args.precision)
                acfg.set_state_dicts(vs=syn.make_vs_state_dict(0, "default"), vd=syn.make_vd_state_dict(1))
                walls = []
                for _ in range(4):
                    t0 = time.perf_counter()
                    df_dyn, df_stat = gpv.preprocess_video_and_predict(path_images=os.path.join(td, "clip"), save_path=td, fps=25,
                                                                       total_frames=1500)
                    walls.append(time.perf_counter() - t0)
                acfg.reset()
                micro["dropin_video_1500_jpeg_files"] = {"frames_per_s_wall": 1500 / min(walls[1:]), "ms_wall": min(walls[1:]) * 1e3,
                                                         "first_call_ms": walls[0] * 1e3, "rows": int(len(df_stat)),
                                                         "what": "get_prob_video.preprocess_video_and_predict on 1500 JPEG crops (160 - 319 px "
                                                                 "a side, every file its own size) of one clip on disk (page cache), wall clock of the "
                                                                 "whole call"}
        except Exception as e:  # pragma: no cover
            micro["dropin_video_1500_jpeg_files"] = {"skipped": f"{type(e).__name__}: {e}"}
        # the face detector that produces the crops (RetinaFace-ResNet50 on raw 1080p frames;
        # the reference: one batch-1 call per frame, data/get_face_images.py:45-61)
        try:
            from avcer_b200 import nets as anets

            fnet = anets.RetinaFaceNet(syn.make_retinaface_state_dict(5, "spread"), args.precision, str(dev))
            fframes = torch.from_numpy(syn.make_frames(7, 8, 1080, 1920)).to(dev)
            fl2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            fnet.detect(fframes)
            ops.PROFILE = fprof = []
            fnet.detect(fframes)
            torch.cuda.synchronize()
            ops.PROFILE = None
            fflop = sum(p[1] for p in fprof)
            fts = []
            for _ in range(5):
                fl2.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fnet.detect(fframes)
                b.record()
                torch.cuda.synchronize()
                fts.append(a.elapsed_time(b))
            ft = statistics.median(fts)
            micro["face_detect_8x1080p"] = {"bound": "tensor", "frames_per_s": 8 / ft * 1e3, "ms": ft, "gflop_per_frame": fflop / 8 / 1e9,
                                            "achieved": fflop / ft / 1e9, "unit": "TFLOP/s", "peak": peaks["tf_burst"],
                                            "frac": fflop / ft / 1e9 / peaks["tf_burst"],
                                            "parity": "detections, track ids and crop files equal the unmodified reference's (tests/test_gpu_face.py)"}
            del fnet, fframes, fl2
        except Exception as e:  # pragma: no cover
            micro["face_detect_8x1080p"] = {"skipped": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        cpu = cpu_baseline()
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic", "config": workload,
            "audio_seconds_per_sec": sum(durations) * args.steps / (ms / 1e3),
            "shard_frames": runner.counts, "collective": {"op": "all_gather_into_tensor", "bytes_per_rank": int(runner.send.numel() * 4),
                                                         "layout": "per-frame VS probabilities [n,7] + VD logits [n,7] + audio mean logits [n,8], fp32"},
            "roofline": roofline, "kernels": dict(extra, **micro), "vs_resnet50_b256": vs_alone, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_isolated_ms,
            "host_wall_ms_per_step_in_timed_loop": host_enqueue_ms, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
