// tcgen05 implicit-GEMM kernel for sm_100a: one kernel serves every dense contraction on the
// hot path -- ResNet-50 1x1 / strided 1x1 / 3x3 "same" / 7x7 stem convolutions (reference:
// src/architectures/video.py:13-35,98-100,141-148), the wav2vec2 conv1d feature extractor,
// grouped positional conv and all Linear layers (audio_8_cl.py:131-190, attention_layers.py:80-144).
//
//   D[row, co] = act( sum_{tap, c} A[row @ tap, c] * Wt[co, tap*Cin + c] + bias[co] (+ residual[row, co]) )
//
// * A (activations, channels-last bf16) is fetched by TMA through a rank-5 tiled tensor map
//   (c, w, h, n, t).  An M tile is a *box* of bw*bh*bn <= 128 output positions; filter taps are
//   coordinate offsets, and "same" zero padding is TMA out-of-bounds zero fill.
// * Wt (weights, [Cout, taps*Cin] bf16, K contiguous) is fetched by a rank-2 TMA.
// * Both land in shared memory in the 128B (or 64B) swizzled K-major layout that
//   tcgen05.mma consumes directly; accumulators live in TMEM (2 x BN fp32 columns, double
//   buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
// * Epilogue (bf16 output): the residual tile is prefetched by TMA into shared memory while the
//   MMAs run; 4 warps read the accumulator with tcgen05.ld, add bias / residual, apply the
//   activation, and write the bf16 tile into a 128B-swizzled staging buffer (bank-conflict free);
//   one thread hands it to a TMA store (cp.async.bulk.tensor ... bulk_group), which clips partial
//   boxes at the tensor border -- no per-row masks, every global transaction is a full line.
// * Warp roles: warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM
//   allocator, warps4-11 = epilogue (two warps per TMEM lane quarter, interleaved column chunks).
// * Persistent: grid = min(#tiles, #SM); tiles are strided across CTAs, N fastest so CTAs of
//   one wave share the same activation box in L2.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace avcer {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
// Output modes of the kernel template
enum OutMode : int { OUT_TMA = 0, OUT_TMA_RES = 1, OUT_DIRECT_F32 = 2 };

struct TcGemmParams {
  int num_tiles, tiles_n;
  int reverse;             // 1: the persistent CTAs walk the tiles from the last to the first (see avcer_contract_desc.reverse_tiles)
  int bw, bh, bn;          // box extents of one M tile (bw*bh*bn <= 128)
  int tw, th, tn;          // tiles along w / h / n
  int W, H, NB;            // valid output extents (direct-store epilogue mask)
  int taps_w, taps_h;      // filter taps
  int off_w, off_h;        // coordinate offset of tap (0,0)  (= -padding)
  int a_step;              // spatial step of the A box (>= 1): output pixel w reads input pixel a_step*w + off_w + tap
  int tap_h_in_dim4;       // 1: the h-tap index is coordinate 4 of the A map (stem: strided rows)
  int kchunks;             // K chunks (of BK) per tap
  int a_c0_per_ntile;      // channel-coordinate shift per N tile (grouped conv), else 0
  int a_strip;             // 1: A box is an un-swizzled contiguous strip whose rows overlap (row r starts 16 B after
                           //    row r-1): the im2col of a strided conv row, expressed in the UMMA descriptor alone
  unsigned a_bytes;        // bytes delivered by one A TMA load
  const void* b_packed;    // strip mode: weights pre-packed per K chunk in UMMA no-swizzle core-matrix order
                           //   [k/8][n/8][8 rows][8 elems]; one 1-D bulk copy per chunk instead of a 2-D TMA box
  int Cout;
  long long out_sw, out_sh, out_sn;   // output element strides of (w, h, n)   (direct-store mode)
  const float* bias;                  // [Cout] or nullptr
  void* out;                          // fp32 output (direct-store mode)
  int act;
  int res_after_act;     // 0: act(acc + bias + res)   1: act(acc + bias) + res
  unsigned long long* trace;   // development aid (avcer_debug_set_trace): per-tile clock64 stamps of the first CTAs, else nullptr
};

constexpr int kTraceCtas = 4, kTraceTiles = 64, kTraceSlots = 16;
__device__ __forceinline__ void trace_stamp(const TcGemmParams& p, int local_tile, int slot) {
  if (p.trace != nullptr && blockIdx.x < kTraceCtas && local_tile < kTraceTiles)
    p.trace[(blockIdx.x * kTraceTiles + local_tile) * kTraceSlots + slot] = clock64();
}

template <int BN, int BK, int MODE, int KSUB, int OCC>
struct TcGemmCfg {
  static constexpr int BM = 128;
  static constexpr int A_STAGE = BM * BK * 2;
  static constexpr int B_STAGE = BN * BK * 2;
  static constexpr int SUB = A_STAGE + B_STAGE;                                 // one K chunk (BK) of A and B
  static constexpr int STAGE = KSUB * SUB;                                      // one pipeline stage = KSUB chunks
  static constexpr int HALF = BM * 128;                                        // 128 rows x 64 bf16 columns, 128B-swizzled
  // OCC = CTAs per SM.  OCC 2 (small tiles): two independent pipelines share an SM, so the epilogue of one
  // CTA overlaps the loads / MMAs of the other; 4 epilogue warps, single staging slots, half the budget.
  static constexpr int EPI_WARPS = (OCC == 2) ? 4 : 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int C_SLOTS = (MODE == OUT_DIRECT_F32) ? 0 : (OCC == 2 ? 1 : 2);   // output staging ring (64-column halves)
  static constexpr int R_SLOTS = (MODE == OUT_TMA_RES) ? (OCC == 2 ? 1 : 2) : 0;      // residual prefetch ring
  static constexpr int BUDGET = (OCC == 2) ? 111 * 1024 : 224 * 1024;
  static constexpr int FIT = (BUDGET - (C_SLOTS + R_SLOTS) * HALF) / STAGE;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = STAGES * STAGE + (C_SLOTS + R_SLOTS) * HALF + 1024 /*align slack*/;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int HALVES = BN / 64;
  static constexpr uint32_t LAYOUT = (BK == 64) ? 2u : 4u;      // SWIZZLE_128B : SWIZZLE_64B
  static constexpr uint32_t SBO = 8u * BK * 2u;                 // 8-row group pitch
  static_assert(BK == 64 || BK == 32, "BK must be one swizzle row");
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN must be 64, 128 or 256");
  static_assert(OCC == 1 || TMEM_COLS <= 256, "two CTAs per SM share the 512 TMEM columns");
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <int BN, int BK, int MODE, int KSUB, int OCC>
__global__ void __launch_bounds__(128 + ((OCC == 2) ? 128 : 256), OCC)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const TcGemmParams p) {
  using Cfg = TcGemmCfg<BN, BK, MODE, KSUB, OCC>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::STAGES + 8];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // stage s, chunk j: A at a_base + (s*KSUB + j)*A_STAGE, B at b_base + (s*KSUB + j)*B_STAGE
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + Cfg::STAGES * KSUB * Cfg::A_STAGE;
  const uint32_t c_base = smem_base + Cfg::STAGES * Cfg::STAGE;
  const uint32_t r_base = c_base + Cfg::C_SLOTS * Cfg::HALF;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  auto rfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + a); };
  auto rfree_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 6 + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = p.bw * p.bh * p.bn;
  const int k_iters = p.taps_w * p.taps_h * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE != OUT_DIRECT_F32) tma_prefetch_desc(&tmC);
    if (MODE == OUT_TMA_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::EPI_WARPS);   // one arrive per epilogue warp
      mbar_init(rfull_bar(a), 1);
      mbar_init(rfree_bar(a), Cfg::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_slot_s), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);
  pdl_wait();                  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = p.a_bytes + Cfg::B_STAGE;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
        const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
        const int nt = te % p.tiles_n;
        const int mt = te / p.tiles_n;
        const int w0 = (mt % p.tw) * p.bw;
        const int h0 = ((mt / p.tw) % p.th) * p.bh;
        const int n0 = (mt / (p.tw * p.th)) * p.bn;
        const int c_shift = nt * p.a_c0_per_ntile;
        // K iterations are (ty, tx, kc) flattened; one stage carries up to KSUB of them under one barrier so
        // that the single-thread MMA issue loop pays its wait/commit latency once per KSUB chunks
        for (int it = 0; it < k_iters; it += KSUB) {
          const int nsub = (k_iters - it) < KSUB ? (k_iters - it) : KSUB;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), static_cast<uint32_t>(nsub) * (p.a_bytes + Cfg::B_STAGE));
          for (int j = 0; j < nsub; ++j) {
            const int kk = it + j;
            const int kc = kk % p.kchunks;
            const int tap = kk / p.kchunks;
            const int tx = tap % p.taps_w, ty = tap / p.taps_w;
            tma_load_5d(a_base + (stage * KSUB + j) * Cfg::A_STAGE, &tmA, full_bar(stage), p.a_strip ? 0 : kc * BK + c_shift,
                        p.a_strip ? 0 : w0 * p.a_step + p.off_w + tx, h0 * p.a_step + p.off_h + (p.tap_h_in_dim4 ? 0 : ty), n0,
                        p.tap_h_in_dim4 ? ty : 0);
            if (p.b_packed != nullptr)
              bulk_load_1d(b_base + (stage * KSUB + j) * Cfg::B_STAGE,
                           static_cast<const uint8_t*>(p.b_packed) + (size_t)(nt * k_iters + kk) * Cfg::B_STAGE, Cfg::B_STAGE,
                           full_bar(stage));
            else
              tma_load_2d(b_base + (stage * KSUB + j) * Cfg::B_STAGE, &tmB, full_bar(stage), kk * BK, nt * BN);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; it += KSUB) {
          const int nsub = (k_iters - it) < KSUB ? (k_iters - it) : KSUB;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          for (int j = 0; j < nsub; ++j) {
            const uint32_t a_addr = a_base + (stage * KSUB + j) * Cfg::A_STAGE;
            // strip mode: SWIZZLE_NONE K-major core matrices (8 rows x 16 B, rows 16 B apart), LBO = 16 B, SBO = 128 B
            const uint64_t adesc = p.a_strip ? umma_desc_nosw(a_addr, 16u, 128u) : umma_desc_kmajor(a_addr, Cfg::SBO, Cfg::LAYOUT);
            const uint32_t b_addr = b_base + (stage * KSUB + j) * Cfg::B_STAGE;
            // packed B: core matrices [k/8][n/8]: LBO (next k core) = BN/8 * 128 B, SBO (next n core) = 128 B
            const uint64_t bdesc = p.b_packed != nullptr ? umma_desc_nosw(b_addr, BN * 16u, 128u)
                                                         : umma_desc_kmajor(b_addr, Cfg::SBO, Cfg::LAYOUT);
            const uint32_t b_step = p.b_packed != nullptr ? (2u * BN * 16u) >> 4 : 2u;   // K=16 = two k cores
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // swizzled operands: +32 B per K=16 step inside the swizzle row (start-address field += 2)
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + b_step * k, idesc, (it | j | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ residual producer (own warp: never blocks A/B loads)
    if (MODE == OUT_TMA_RES && lane == 0) {
      uint32_t hcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
        const int nt = te % p.tiles_n;
        const int mt = te / p.tiles_n;
        const int w0 = (mt % p.tw) * p.bw;
        const int h0 = ((mt / p.tw) % p.th) * p.bh;
        const int n0 = (mt / (p.tw * p.th)) * p.bn;
        for (int hf = 0; hf < Cfg::HALVES; ++hf, ++hcount) {
          constexpr int RS = Cfg::R_SLOTS > 0 ? Cfg::R_SLOTS : 1;
          const int slot = hcount % RS;
          mbar_wait(rfree_bar(slot), ((hcount / RS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(rfull_bar(slot), static_cast<uint32_t>(rows) * 128);
          tma_load_5d(r_base + slot * Cfg::HALF, &tmR, rfull_bar(slot), nt * BN + hf * 64, w0, h0, n0, 0);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;           // OCC 1: two warps share a quarter (even / odd 32-column chunks)
    constexpr int EPI_THREADS = Cfg::EPI_WARPS * 32;
    constexpr int CPW = 8 / Cfg::EPI_WARPS;    // 32-column chunks of a 64-column half handled by one warp
    const int r = q * 32 + lane;               // row of the tile owned by this thread
    const bool store_thread = (threadIdx.x == 128);
    uint32_t hcount = 0;                       // 64-column halves handled so far (ring positions)
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1u;
      const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
      const int nt = te % p.tiles_n;
      const int mt = te / p.tiles_n;
      const int w0 = (mt % p.tw) * p.bw;
      const int h0 = ((mt / p.tw) % p.th) * p.bh;
      const int n0 = (mt / (p.tw * p.th)) * p.bn;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();

      auto load_chunk = [&](int c, float (&f)[32]) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        const int co = nt * BN + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr && co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co + j));
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
      };
      auto apply_act = [&](float (&f)[32]) {
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] < 0.0f ? 0.0f : f[j];   // NaN-propagating like torch.relu
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            if (MODE == OUT_DIRECT_F32) { f[j] = gelu_erf(f[j]); f[j + 1] = gelu_erf(f[j + 1]); }
            else gelu_erf_fast2(f[j], f[j + 1]);
          }
        }
      };

      if (MODE == OUT_DIRECT_F32) {
        const int dw = r % p.bw, dh = (r / p.bw) % p.bh, dn = r / (p.bw * p.bh);
        const int w = w0 + dw, h = h0 + dh, n = n0 + dn;
        const bool valid = (r < rows) && (w < p.W) && (h < p.H) && (n < p.NB);
#pragma unroll 1
        for (int c = grp; c < BN / 32; c += 2 / CPW) {
          float f[32];
          load_chunk(c, f);
          apply_act(f);
          const int co = nt * BN + c * 32;
          if (valid && co < p.Cout) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + w * p.out_sw + h * p.out_sh + n * p.out_sn + co);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      } else {
#pragma unroll 1
        for (int hf = 0; hf < Cfg::HALVES; ++hf, ++hcount) {
          constexpr int CS = Cfg::C_SLOTS > 0 ? Cfg::C_SLOTS : 1;
          constexpr int RS = Cfg::R_SLOTS > 0 ? Cfg::R_SLOTS : 1;
          const int slot = hcount % CS;
          const int rslot = hcount % RS;
          // staging slot `slot` was last read by the TMA store issued CS halves ago
          if (store_thread) bulk_wait_group_read<CS - 1>();
          named_bar_sync(1, EPI_THREADS);
          const uint32_t cbuf = c_base + slot * Cfg::HALF;
          const uint32_t rbuf = r_base + rslot * Cfg::HALF;
          const uint32_t row_off = r * 128;
          if (MODE == OUT_TMA_RES) mbar_wait(rfull_bar(rslot), (hcount / RS) & 1u);
#pragma unroll
          for (int cc = 0; cc < CPW; ++cc) {
            const int sub = (CPW == 1) ? grp : cc;   // which 32-column chunk of this half
            float f[32];
            load_chunk(2 * hf + sub, f);
            const int j0 = sub * 4;                  // first 16-byte chunk of this thread inside the 128 B row
            if (MODE == OUT_TMA_RES) {
              if (p.res_after_act) apply_act(f);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                ld_shared_v4(rbuf + row_off + (((j0 + j) ^ (r & 7)) << 4), u);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t = __bfloat1622float2(h2[e]);
                  f[j * 8 + e * 2] += t.x;
                  f[j * 8 + e * 2 + 1] += t.y;
                }
              }
              if (!p.res_after_act) apply_act(f);
            } else {
              apply_act(f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
              st_shared_v4(cbuf + row_off + (((j0 + j) ^ (r & 7)) << 4), u);
            }
          }
          fence_proxy_async();                 // generic-proxy smem writes -> visible to the TMA engine; residual reads
                                               // ordered before the TMA refill of the slot
          if (MODE == OUT_TMA_RES) {
            __syncwarp();
            if (lane == 0) mbar_arrive(rfree_bar(rslot));
          }
          named_bar_sync(2, EPI_THREADS);
          if (store_thread) {
            if (nt * BN + hf * 64 < p.Cout) tma_store_5d(&tmC, cbuf, nt * BN + hf * 64, w0, h0, n0, 0);
            bulk_commit_group();
          }
        }
      }
      // accumulator is free again
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (MODE != OUT_DIRECT_F32 && store_thread) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
