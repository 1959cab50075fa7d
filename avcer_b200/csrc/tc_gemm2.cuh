// Two-SM (cta_group::2) flavour of the tcgen05 implicit-GEMM kernel: a cluster of two CTAs on one TPC
// computes a 256 x 256 output tile with UMMA M = 256.  Each CTA stages its own 128 activation rows and
// HALF of the 256-row weight tile, so the shared-memory operand traffic per MMA is 64 B/clk per SM instead
// of 96 (128x256 single-CTA) or 128 (128x128) -- the limiter of the single-CTA tiles.
//
// Protocol (leader = cluster rank 0):
//   * both CTAs' TMA producers load into their own shared memory but complete their bytes on the LEADER's
//     `full` barrier (cta_group::2 TMA, barrier address with the peer bit cleared);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (A rows 0-127 / B half 0 from CTA 0,
//     rows 128-255 / half 1 from CTA 1; accumulator rows land in each CTA's own TMEM) and multicasts its
//     commits to the `empty` / `tfull` barriers of both CTAs;
//   * each CTA runs its own epilogue on its 128 rows (same ring-staged TMA-store epilogue as tc_gemm.cuh);
//     all epilogue warps of both CTAs release the accumulator on the leader's `tempty` barrier.
// Geometry (boxes, taps, padding by TMA zero fill, persistent scheduling) is identical to tc_gemm.cuh.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's barrier at this offset
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int MODE, int BN_, int RS_ = 4, int FLAT_ = 0>
struct TcGemm2Cfg {
  static constexpr int BN = BN_;                 // N of the pair tile (each CTA stages BN/2 weight rows): 256 or 128
  static constexpr int BK = 64;
  static constexpr int A_STAGE = 128 * BK * 2;   // own 128 activation rows
  static constexpr int B_STAGE = (BN / 2) * BK * 2;   // own half of the weight tile
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int HALF = 128 * 128;
  static constexpr int C_SLOTS = 2;              // output staging ring (a 4-slot ring measured no faster)
  static constexpr int R_SLOTS = (MODE == OUT_TMA_RES) ? RS_ : 0;   // residual prefetch ring: 64 KB of residual reads in flight per SM
  static constexpr int R_RING = R_SLOTS > 0 ? R_SLOTS : 1;
  static constexpr int BUDGET = 224 * 1024;
  // FLAT epilogue (outputs that are plain [M, Cout] rows): every epilogue warp owns two 32-row x 128 B staging
  // slabs (and two residual slabs) and moves them with its own TMA operations -- no CTA-wide barriers.
  static constexpr int SLAB = 32 * 128;
  static constexpr int C_BYTES = FLAT_ ? 8 * 2 * SLAB : C_SLOTS * HALF;
  static constexpr int R_BYTES = (MODE == OUT_TMA_RES) ? (FLAT_ ? 8 * 2 * SLAB : R_SLOTS * HALF) : 0;
  static constexpr int N_RBARS = FLAT_ ? 16 : 2 * R_RING;        // FLAT: one per (warp, slab); else rfull + rfree rings
  static constexpr int FIT = (BUDGET - C_BYTES - R_BYTES) / STAGE;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM = STAGES * STAGE + C_BYTES + R_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int HALVES = BN / 64;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
};

template <int MODE, int BN_, int RS_, int FLAT_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                const TcGemmParams p) {
  using Cfg = TcGemm2Cfg<MODE, BN_, RS_, FLAT_>;
  constexpr int BN = Cfg::BN, BK = Cfg::BK;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::STAGES + 4 + Cfg::N_RBARS];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + Cfg::STAGES * Cfg::A_STAGE;
  const uint32_t c_base = smem_base + Cfg::STAGES * Cfg::STAGE;
  const uint32_t r_base = c_base + Cfg::C_BYTES;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  auto rfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + a); };
  auto rfree_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + Cfg::R_RING + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int rows = p.bw * p.bh * p.bn;
  const int k_iters = p.taps_w * p.taps_h * p.kchunks;
  const int m_tiles = p.tw * p.th * p.tn;
  const int pair_tiles = ((m_tiles + 1) >> 1) * p.tiles_n;   // tiles_n counts BN-wide N tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (MODE == OUT_TMA_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);        // leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(empty_bar(s), 1);       // multicast commit from the leader's MMA thread
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * Cfg::EPI_WARPS);   // epilogue warps of BOTH CTAs (used on the leader only)
    }
    if (FLAT_) {
      for (int a = 0; a < Cfg::N_RBARS; ++a) mbar_init(rfull_bar(a), 1);      // (epilogue warp, slab) residual barriers
    } else {
      for (int a = 0; a < Cfg::R_RING; ++a) {
        mbar_init(rfull_bar(a), 1);
        mbar_init(rfree_bar(a), Cfg::EPI_WARPS);
      }
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(smem_u32(&tmem_slot_s), Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                   // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);
  pdl_wait();                  // the prologue above overlapped the previous kernel's tail
  pdl_launch_dependents();

  auto tile_coords = [&](int pt, int& nt, int& w0, int& h0, int& n0) {
    if (p.reverse) pt = pair_tiles - 1 - pt;
    nt = pt % p.tiles_n;
    const int mt = 2 * (pt / p.tiles_n) + (int)rank;       // phantom tile when m_tiles is odd: n0 >= NB -> all out of bounds
    w0 = (mt % p.tw) * p.bw;
    h0 = ((mt / p.tw) % p.th) * p.bh;
    n0 = (mt / (p.tw * p.th)) * p.bn;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_pair = 2u * (p.a_bytes + Cfg::B_STAGE);
      int local = 0;
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++local) {
        int nt, w0, h0, n0;
        tile_coords(pt, nt, w0, h0, n0);
        int kcol = 0;
        trace_stamp(p, local, 0);
        for (int ty = 0; ty < p.taps_h; ++ty) {
          for (int tx = 0; tx < p.taps_w; ++tx) {
            for (int kc = 0; kc < p.kchunks; ++kc, kcol += BK) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), tx_pair);
              tma_load_5d_2sm(a_base + stage * Cfg::A_STAGE, &tmA, full_bar(stage), kc * BK, w0 * p.a_step + p.off_w + tx,
                              h0 * p.a_step + p.off_h + (p.tap_h_in_dim4 ? 0 : ty), n0, p.tap_h_in_dim4 ? ty : 0);
              tma_load_2d_2sm(b_base + stage * Cfg::B_STAGE, &tmB, full_bar(stage), kcol, nt * BN + (int)rank * (BN / 2));
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
        trace_stamp(p, local, 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1u;
        trace_stamp(p, local, 2);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        trace_stamp(p, local, 3);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (it == 0) trace_stamp(p, local, 4);
          const uint64_t adesc = umma_desc_kmajor(a_base + stage * Cfg::A_STAGE, 1024u, 2u);
          const uint64_t bdesc = umma_desc_kmajor(b_base + stage * Cfg::B_STAGE, 1024u, 2u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (it | k) != 0 ? 1u : 0u);
          umma_commit_2sm(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(tfull_bar(acc));
        trace_stamp(p, local, 5);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ residual producer (per CTA, own rows)
    if (MODE == OUT_TMA_RES && !FLAT_ && lane == 0) {
      uint32_t hcount = 0;
      int local = 0;
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++local) {
        int nt, w0, h0, n0;
        tile_coords(pt, nt, w0, h0, n0);
        for (int hf = 0; hf < Cfg::HALVES; ++hf, ++hcount) {
          if (hf == Cfg::HALVES - 1) trace_stamp(p, local, 13);
          const int slot = hcount % Cfg::R_RING;
          mbar_wait(rfree_bar(slot), ((hcount / Cfg::R_RING) & 1u) ^ 1u);
          mbar_arrive_expect_tx(rfull_bar(slot), static_cast<uint32_t>(rows) * 128);
          tma_load_5d(r_base + slot * Cfg::HALF, &tmR, rfull_bar(slot), nt * BN + hf * 64, w0, h0, n0, 0);
        }
      }
    }
  } else if (warp >= 4 && FLAT_) {
    // ------------------------------------------------------------ FLAT epilogue: warps run independently.
    // Warp (q, grp) owns rows 32q..32q+31 of the CTA's 128 and the 64-column halves hf = grp, grp+2, ...:
    // tcgen05.ld 64 columns -> bias / residual / activation -> bf16 into its own swizzled 4 KB slab -> its own
    // TMA store (box 64 x 32 rows, clipped at M); the residual slab of its item j+2 is prefetched by the same warp
    // as soon as item j has consumed that slab.  The only cross-warp synchronisation left is tfull / tempty.
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int ew = warp - 4;
    constexpr int HPW = Cfg::HALVES / 2;                     // halves per warp per tile
    const uint32_t cslab = c_base + ew * 2 * Cfg::SLAB;
    const uint32_t rslab = r_base + ew * 2 * Cfg::SLAB;
    const uint32_t rbar = rfull_bar(2 * ew);
    const int my_tiles = cluster_id < pair_tiles ? (pair_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
    const int n_items = my_tiles * HPW;
    auto item_coords = [&](int j, int& col0, int& row0) {
      int pt = cluster_id + (j / HPW) * num_clusters;
      if (p.reverse) pt = pair_tiles - 1 - pt;
      col0 = (pt % p.tiles_n) * BN + (grp + 2 * (j % HPW)) * 64;
      row0 = (2 * (pt / p.tiles_n) + (int)rank) * 128 + q * 32;
    };
    auto issue_res = [&](int j) {
      if (j < n_items) {
        int col0, row0;
        item_coords(j, col0, row0);
        mbar_arrive_expect_tx(rbar + 8u * (j & 1), Cfg::SLAB);
        tma_load_5d(rslab + (j & 1) * Cfg::SLAB, &tmR, rbar + 8u * (j & 1), col0, row0, 0, 0, 0);
      }
    };
    if (MODE == OUT_TMA_RES && lane == 0) { issue_res(0); issue_res(1); }
    const uint32_t row_off = lane * 128;
    int j = 0;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1u;
      if (threadIdx.x == 128) trace_stamp(p, local, 6);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (threadIdx.x == 128) trace_stamp(p, local, 7);
#pragma unroll 1
      for (int h = 0; h < HPW; ++h, ++j) {
        int col0, row0;
        item_coords(j, col0, row0);
        const int hf = grp + 2 * h;
        const uint32_t cbuf = cslab + (j & 1) * Cfg::SLAB;
        const uint32_t rbuf = rslab + (j & 1) * Cfg::SLAB;
        if (lane == 0) bulk_wait_group_read<1>();            // the store that last read this slab (item j-2) is done with it
        __syncwarp();
        if (threadIdx.x == 128 && h == 0) trace_stamp(p, local, 8);
        uint32_t v[2][32];
        const uint32_t taddr = tmem_base + acc * BN + hf * 64 + (static_cast<uint32_t>(q * 32) << 16);
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_32x32(taddr + 32, v[1]);
        if (MODE == OUT_TMA_RES) mbar_wait(rbar + 8u * (j & 1), (j >> 1) & 1u);
        if (threadIdx.x == 128 && h == 0) trace_stamp(p, local, 9);
        tmem_ld_wait();
        if (threadIdx.x == 128 && h == 0) trace_stamp(p, local, 10);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[cc][i]);
          const int co = col0 + cc * 32;
          if (p.bias != nullptr && co < p.Cout) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co + i));
              add_f32x2(f[i], f[i + 1], b.x, b.y);
              add_f32x2(f[i + 2], f[i + 3], b.z, b.w);
            }
          }
          if (p.act != ACT_GELU && !p.res_after_act) {
            // ReLU / identity: ~3 instructions per element instead of ~6 -- packed fp32 adds, the residual unpacked with
            // one shift and one mask per bf16 pair, ReLU as one NaN-propagating packed max on the rounded pair
            // (rounding is monotonic and keeps zeros and NaNs, so relu(round(x)) == round(relu(x)))
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (MODE == OUT_TMA_RES) {
                uint4 u;
                ld_shared_v4(rbuf + row_off + (((cc * 4 + i) ^ (lane & 7)) << 4), u);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#ifdef AVCER_HALF
                  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                  add_f32x2(f[i * 8 + e * 2], f[i * 8 + e * 2 + 1], t.x, t.y);
#else
                  add_f32x2(f[i * 8 + e * 2], f[i * 8 + e * 2 + 1], __uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
#endif
                }
              }
              uint4 o;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h2[e] = __floats2bfloat162_rn(f[i * 8 + e * 2], f[i * 8 + e * 2 + 1]);
                if (p.act == ACT_RELU) h2[e] = __hmax2_nan(h2[e], zero2);
              }
              st_shared_v4(cbuf + row_off + (((cc * 4 + i) ^ (lane & 7)) << 4), o);
            }
          } else {
            auto apply_act = [&]() {
              if (p.act == ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = f[i] < 0.0f ? 0.0f : f[i];
              } else if (p.act == ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) gelu_erf_fast2(f[i], f[i + 1]);
              }
            };
            if (MODE == OUT_TMA_RES) {
              if (p.res_after_act) apply_act();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u;
                ld_shared_v4(rbuf + row_off + (((cc * 4 + i) ^ (lane & 7)) << 4), u);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t = __bfloat1622float2(h2[e]);
                  f[i * 8 + e * 2] += t.x;
                  f[i * 8 + e * 2 + 1] += t.y;
                }
              }
              if (!p.res_after_act) apply_act();
            } else {
              apply_act();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[i * 8 + e * 2], f[i * 8 + e * 2 + 1]);
              st_shared_v4(cbuf + row_off + (((cc * 4 + i) ^ (lane & 7)) << 4), u);
            }
          }
        }
        if (threadIdx.x == 128 && h == 0) trace_stamp(p, local, 11);
        fence_proxy_async();
        __syncwarp();
        if (threadIdx.x == 128 && h == 0) trace_stamp(p, local, 12);
        if (lane == 0) {
          if (col0 < p.Cout) tma_store_5d(&tmC, cbuf, col0, row0, 0, 0, 0);
          bulk_commit_group();
          if (MODE == OUT_TMA_RES) issue_res(j + 2);         // every lane has consumed this residual slab (syncwarp above)
        }
        if (threadIdx.x == 128 && h < 2) trace_stamp(p, local, 13 + h);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (threadIdx.x == 128) trace_stamp(p, local, 15);
    }
    if (lane == 0) bulk_wait_group<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (per CTA, own 128 rows)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const bool store_thread = (threadIdx.x == 128);
    uint32_t hcount = 0;
    int local = 0;
    for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1u;
      int nt, w0, h0, n0;
      tile_coords(pt, nt, w0, h0, n0);
      if (store_thread) trace_stamp(p, local, 6);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (store_thread) trace_stamp(p, local, 7);
#pragma unroll 1
      for (int hf = 0; hf < Cfg::HALVES; ++hf, ++hcount) {
        const int slot = hcount % Cfg::C_SLOTS;
        const int rslot = hcount % Cfg::R_RING;
        if (store_thread) bulk_wait_group_read<Cfg::C_SLOTS - 1>();
        named_bar_sync(1, 256);
        const uint32_t cbuf = c_base + slot * Cfg::HALF;
        const uint32_t rbuf = r_base + rslot * Cfg::HALF;
        const uint32_t row_off = r * 128;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + (2 * hf + grp) * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        const int co = nt * BN + (2 * hf + grp) * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr && co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co + j));
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
        auto apply_act = [&]() {
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] < 0.0f ? 0.0f : f[j];
          } else if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_erf_fast2(f[j], f[j + 1]);
          }
        };
        const int j0 = grp * 4;
        if (MODE == OUT_TMA_RES) {
          if (p.res_after_act) apply_act();
          mbar_wait(rfull_bar(rslot), (hcount / Cfg::R_RING) & 1u);
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
          fence_proxy_async();                 // generic-proxy reads of the slot ordered before the TMA refill (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(rfree_bar(rslot));
          if (!p.res_after_act) apply_act();
        } else {
          apply_act();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
          st_shared_v4(cbuf + row_off + (((j0 + j) ^ (r & 7)) << 4), u);
        }
        fence_proxy_async();
        named_bar_sync(2, 256);
        if (store_thread) {
          if (nt * BN + hf * 64 < p.Cout) tma_store_5d(&tmC, cbuf, nt * BN + hf * 64, w0, h0, n0, 0);
          bulk_commit_group();
          if (hf < 4) trace_stamp(p, local, 8 + hf);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (store_thread) trace_stamp(p, local, 12);    // both CTAs release the accumulator on the leader
    }
    if (store_thread) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the peer may still multicast into / arrive on this CTA's barriers until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
