// tcgen05 self-attention for short sequences (T <= 208 keys, head dim 64), bf16 in / fp32 accumulate / bf16 out:
// softmax(Q K^T * scale) V per (window, head) of HF Wav2Vec2Attention (12 encoder layers, 16 heads x 64) and of
// src/architectures/attention_layers.py:10-38 (tl2).  Replaces the mma.sync kernel of attention_tc.cu for dh = 64
// (that kernel ran the legacy tensor path at 25 % pipe-active, latency-bound, 14 % of the audio forward).
//
// Work unit = one (window, head); its one or two 128-query-row tiles are handled side by side by two softmax warp
// groups (group g = row tile g = TMEM buffer S[g] = P buffer g).  Per tile
//   S = Q K^T : 4 x tcgen05.mma 128 x Npad x 16 (Q tile and K rows are K-major SWIZZLE_128B TMA boxes of the packed
//               qkv rows; Npad = T rounded up to 16), accumulator S[g] in TMEM
//   softmax   : thread = query row; pass 1 reads S for the row maximum, pass 2 re-reads it, p = exp2((s - max) * scale')
//               with keys >= T masked, row sum in fp32, P written as bf16 into shared memory in the K-major swizzled
//               layout of an A operand, signalled chunk by chunk (64 keys) so P V starts under the rest of pass 2
//   O = P V   : Npad/16 x tcgen05.mma 128 x 64 x 16 with V as an MN-MAJOR B operand -- V rows [key][64] are exactly
//               the canonical MN-major SWIZZLE_128B atom (8 keys x 128 B), so no transpose is ever materialised.  O is
//               accumulated into the first 64 columns of the tile's own S buffer (already consumed by pass 2)
//   epilogue  : O / rowsum -> bf16 -> global (rows < T only)
// Q / K and V are single-buffered but refilled early: Q and K of the next head are requested as soon as both Q K^T
// of this head have completed (before its softmax), V as soon as both P V have.
// Warp roles: warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warps 4-7 / 8-11 softmax groups (TMEM lane
// quarter = warp % 4).  Persistent, one CTA per SM.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

struct Att5Params {
  int n, t, heads;          // windows, tokens per window, heads (head dim 64)
  int npad;                 // keys rounded up to 16
  int mtiles;               // ceil(t / 128): 1 or 2
  int units;                // n * heads
  float scale_log2e;
  __nv_bfloat16* out;       // [n*t, heads*64]
};

struct Att5Cfg {
  static constexpr int MAXT = 208;
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int KV_BYTES = MAXT * 128;
  static constexpr int P_CHUNK = 128 * 128;                    // 128 rows x 64 keys
  static constexpr int P_BYTES = 4 * P_CHUNK;
  static constexpr int SMEM = 2 * Q_BYTES + 2 * KV_BYTES + 2 * P_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;                        // S0 @0, S1 @256 (<= 208 columns each); O = columns 0..63 of the tile's S
  static constexpr int THREADS = 384;
  static_assert(Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0, "operand bases must stay 1 KB aligned");
  static_assert(SMEM <= 227 * 1024, "attention: shared memory budget");
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kind::f16 instruction descriptor with B given MN-major (bit 16): D=f32, A=B=bf16, A K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 16); }

__global__ void __launch_bounds__(384, 1)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Att5Params p) {
  using Cfg = Att5Cfg;
  extern __shared__ uint8_t smem_raw[];
  // barriers: qk_full qk_empty v_full v_empty s_full[2] s_empty[2] o_full[2] p_full[2][4]
  __shared__ __align__(8) uint64_t bars[18];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base;                                   // two 128-row Q tiles
  const uint32_t k_base = q_base + 2 * Cfg::Q_BYTES;
  const uint32_t v_base = k_base + Cfg::KV_BYTES;
  const uint32_t p_base = v_base + Cfg::KV_BYTES;                      // two P buffers of 4 chunks
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t qk_full = bar_base, qk_empty = bar_base + 8, v_full = bar_base + 16, v_empty = bar_base + 24;
  auto s_full = [&](int g) { return bar_base + 8u * (4 + g); };
  auto s_empty = [&](int g) { return bar_base + 8u * (6 + g); };       // S[g] (and the O inside it) may be overwritten
  auto o_full = [&](int g) { return bar_base + 8u * (8 + g); };        // P V of tile g has completed
  auto p_full = [&](int g, int c) { return bar_base + 8u * (10 + g * 4 + c); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 4);            // one arrive per warp of the group
      mbar_init(o_full(g), 1);
      for (int c = 0; c < 4; ++c) mbar_init(p_full(g, c), 4);
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
  pdl_wait();
  pdl_launch_dependents();

  const int my_units = blockIdx.x < p.units ? (p.units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nck = (p.npad + 63) >> 6;                          // P chunks of 64 keys
  const int ksteps = p.npad >> 4;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int u = 0; u < my_units; ++u) {
        const int unit = blockIdx.x + u * gridDim.x;
        const int h = unit % p.heads, b = unit / p.heads;
        mbar_wait(qk_empty, (u & 1u) ^ 1u);
        mbar_arrive_expect_tx(qk_full, p.mtiles * Cfg::Q_BYTES + Cfg::KV_BYTES);
        // rank-3 maps (column, token, window): tokens >= T of a box are out of bounds and arrive as ZEROS, never as the
        // first rows of the next window (a 0 * NaN in P V would otherwise leak an all-NaN window -- the empty tail
        // window of "mean" padding, data/utils.py:74-89 -- into its predecessor in the batch)
        for (int g = 0; g < p.mtiles; ++g) tma_load_3d(q_base + g * Cfg::Q_BYTES, &tmQ, qk_full, h * 64, g * 128, b);
        tma_load_3d(k_base, &tmKV, qk_full, (p.heads + h) * 64, 0, b);
        mbar_wait(v_empty, (u & 1u) ^ 1u);
        mbar_arrive_expect_tx(v_full, Cfg::KV_BYTES);
        tma_load_3d(v_base, &tmKV, v_full, (2 * p.heads + h) * 64, 0, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, p.npad);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
      for (int u = 0; u < my_units; ++u) {
        mbar_wait(qk_full, u & 1u);
        for (int g = 0; g < p.mtiles; ++g) {                   // S[g] = Q_g K^T
          mbar_wait(s_empty(g), (u & 1u) ^ 1u);
          tc_fence_after();
          const uint64_t adesc = umma_desc_kmajor(q_base + g * Cfg::Q_BYTES, 1024u, 2u);
          const uint64_t bdesc = umma_desc_kmajor(k_base, 1024u, 2u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + g * 256, adesc + 2u * k, bdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(s_full(g));
        }
        umma_commit(qk_empty);                                 // Q / K of the next head may be fetched
        mbar_wait(v_full, u & 1u);
        for (int c = 0; c < nck; ++c) {                        // O_g += P_g[:, chunk c] V[chunk c, :]
          const int k1 = (c * 4 + 4 < ksteps) ? c * 4 + 4 : ksteps;
          for (int g = 0; g < p.mtiles; ++g) {
            mbar_wait(p_full(g, c), u & 1u);                   // chunk written; S[g] columns <= 64c+63 consumed by pass 2
            tc_fence_after();
            for (int ks = c * 4; ks < k1; ++ks) {
              const uint64_t adesc = umma_desc_kmajor(p_base + g * Cfg::P_BYTES + c * Cfg::P_CHUNK + (ks & 3) * 32, 1024u, 2u);
              const uint64_t bdesc = umma_desc_kmajor(v_base + ks * 2048, 1024u, 2u);   // MN-major: 16 keys = 2 atoms of 8 rows
              umma_bf16(tmem_base + g * 256, adesc, bdesc, idesc_o, ks != 0 ? 1u : 0u);
            }
            if (c == nck - 1) umma_commit(o_full(g));
          }
        }
        umma_commit(v_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax groups (thread = query row of row tile `grp`)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int row = grp * 128 + r;                             // query row inside the window
    const uint32_t s_addr = tmem_base + grp * 256 + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t pbuf = p_base + grp * Cfg::P_BYTES + r * 128;
    const int nchunks = (p.npad + 31) >> 5;                    // 32-column TMEM reads per row
    // A warp whose 32 rows all lie beyond T (rows 224..255 of a 199-token window) only keeps the barrier protocol going.
    const bool warp_live = grp * 128 + q * 32 < p.t;
    if (grp < p.mtiles) {
      for (int u = 0; u < my_units; ++u) {
        const int unit = blockIdx.x + u * gridDim.x;
        const int h = unit % p.heads, b = unit / p.heads;
        mbar_wait(s_full(grp), u & 1u);
        tc_fence_after();
        // pass 1: row maximum over the valid keys (TMEM loads software-pipelined: chunk c+1 is in flight under chunk c)
        float m = -INFINITY;
        uint32_t va[32], vb[32];
        auto max_chunk = [&](const uint32_t (&v)[32], int c) {
          if (c * 32 + 32 <= p.t) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < p.t) m = fmaxf(m, __uint_as_float(v[j]));
          }
        };
        if (warp_live) {
          tmem_ld_32x32(s_addr, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld_32x32(s_addr + (c + 1) * 32, vb);
            max_chunk(va, c);
            if (c + 1 < nchunks) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld_32x32(s_addr + (c + 2) * 32, va);
              max_chunk(vb, c + 1);
            }
          }
        }
        const float mb = m * p.scale_log2e;
        // pass 2: p = exp2(s * scale' - max * scale'), row sum, bf16 P into the swizzled A-operand chunks
        float l = 0.f;
        auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
          float pr[32];
          if (c * 32 + 32 <= p.t) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pr[j] = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mb));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) pr[j] = (c * 32 + j < p.t) ? ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mb)) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) l += pr[j];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pr[j * 8 + e * 2], pr[j * 8 + e * 2 + 1]);
            st_shared_v4(pbuf + (c >> 1) * Cfg::P_CHUNK + ((((c & 1) * 4 + j) ^ (r & 7)) << 4), w);
          }
        };
        auto chunk_done = [&](int c) {                         // after the second half (or the tail) of a 64-key chunk
          if ((c & 1) == 1 || c == nchunks - 1) {
            tc_fence_before();
            fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(grp, c >> 1));
          }
        };
        if (warp_live) tmem_ld_32x32(s_addr, va);
        for (int c = 0; c < nchunks; c += 2) {
          if (warp_live) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld_32x32(s_addr + (c + 1) * 32, vb);
            exp_chunk(va, c);
          }
          chunk_done(c);
          if (c + 1 < nchunks) {
            if (warp_live) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld_32x32(s_addr + (c + 2) * 32, va);
              exp_chunk(vb, c + 1);
            }
            chunk_done(c + 1);
          }
        }
        // O = P V lands in columns 0..63 of this group's S buffer
        mbar_wait(o_full(grp), u & 1u);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(s_addr, v0);
        tmem_ld_32x32(s_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(grp));              // S / O of this buffer may be overwritten by the next head
        if (row < p.t) {
          const float inv_l = 1.0f / l;
          __nv_bfloat16* orow = p.out + ((long long)b * p.t + row) * (p.heads * 64) + h * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h2[e] = __floats2bfloat162_rn(__uint_as_float(v0[j * 8 + e * 2]) * inv_l, __uint_as_float(v0[j * 8 + e * 2 + 1]) * inv_l);
            reinterpret_cast<uint4*>(orow)[j] = w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h2[e] = __floats2bfloat162_rn(__uint_as_float(v1[j * 8 + e * 2]) * inv_l, __uint_as_float(v1[j * 8 + e * 2 + 1]) * inv_l);
            reinterpret_cast<uint4*>(orow)[4 + j] = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
