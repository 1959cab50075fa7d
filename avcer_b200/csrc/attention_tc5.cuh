// tcgen05 self-attention for short sequences (T <= 208 keys, head dim 64), bf16 in / fp32 accumulate / bf16 out:
// softmax(Q K^T * scale) V per (window, head) of HF Wav2Vec2Attention (12 encoder layers, 16 heads x 64) and of
// src/architectures/attention_layers.py:10-38 (tl2).  Replaces the mma.sync kernel of attention_tc.cu for dh = 64
// (that kernel ran the legacy tensor path at 25 % pipe-active, latency-bound, 14 % of the audio forward).
//
// Work item = (window, head, 128-query-row tile).  Per item
//   S = Q K^T : 4 x tcgen05.mma 128 x Npad x 16 (Q tile and K rows are K-major SWIZZLE_128B TMA boxes of the packed
//               qkv rows; Npad = T rounded up to 16), accumulator S[i & 1] in TMEM
//   softmax   : thread = query row; pass 1 reads S for the row maximum, pass 2 re-reads it, p = exp2((s - max) * scale')
//               with keys >= T masked, row sum in fp32, P written as bf16 into shared memory in the K-major swizzled
//               layout of an A operand, chunk by chunk (64 keys)
//   O = P V   : Npad/16 x tcgen05.mma 128 x 64 x 16 with V as an MN-MAJOR B operand -- V rows [key][64] are exactly
//               the canonical MN-major SWIZZLE_128B atom (8 keys x 128 B), so no transpose is ever materialised.  O is
//               accumulated into the first 64 columns of the item's own S buffer (already consumed by pass 2)
//   epilogue  : O / rowsum -> bf16 -> global (rows < T only)
// Two softmax groups of 4 warps alternate items (group = item parity = S buffer), so the exp work of item i+1 runs
// under the PV / epilogue of item i; the single P buffer is handed over chunk-wise (p_full[c] / p_empty[c]).
// Warp roles: warp0 TMA producer (Q tile + K + V per item, double buffered), warp1 MMA issuer, warp2 TMEM allocator,
// warps 4-7 / 8-11 softmax groups (TMEM lane quarter = warp % 4).  Persistent, one CTA per SM.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

struct Att5Params {
  int n, t, heads;          // windows, tokens per window, heads (head dim 64)
  int npad;                 // keys rounded up to 16
  int mtiles;               // ceil(t / 128)
  int items;                // n * heads * mtiles
  float scale_log2e;
  __nv_bfloat16* out;       // [n*t, heads*64]
};

struct Att5Cfg {
  static constexpr int MAXT = 208;
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int KV_BYTES = MAXT * 128;
  static constexpr int ITEM = Q_BYTES + 2 * KV_BYTES;          // 69 632 B per item buffer
  static constexpr int P_CHUNK = 128 * 128;                    // 128 rows x 64 keys
  static constexpr int P_BYTES = 4 * P_CHUNK;
  static constexpr int SMEM = 2 * ITEM + P_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;                        // S0 @0, S1 @256 (<= 208 columns each); O = columns 0..63 of the item's S
  static constexpr int THREADS = 384;
  static_assert(ITEM % 1024 == 0 && Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0, "operand bases must stay 1 KB aligned");
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
  // barriers: in_full[2] in_empty[2] s_full[2] s_empty[2] o_full[2] p_full[4] p_empty[4]
  __shared__ __align__(8) uint64_t bars[18];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t in_base = smem_base;
  const uint32_t p_base = in_base + 2 * Cfg::ITEM;
  const uint32_t bar_base = smem_u32(bars);
  auto in_full = [&](int b) { return bar_base + 8u * b; };
  auto in_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto s_full = [&](int b) { return bar_base + 8u * (4 + b); };
  auto s_empty = [&](int b) { return bar_base + 8u * (6 + b); };     // S[b] (and the O inside it) may be overwritten
  auto o_full = [&](int b) { return bar_base + 8u * (8 + b); };      // PV of the item in S[b] has completed
  auto p_full = [&](int c) { return bar_base + 8u * (10 + c); };     // P chunk c of the current item is in shared memory
  auto p_empty = [&](int c) { return bar_base + 8u * (14 + c); };    // the MMAs have finished reading P chunk c

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(in_full(b), 1);
      mbar_init(in_empty(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), 4);            // one arrive per warp of the group that owns S[b]
      mbar_init(o_full(b), 1);
    }
    for (int c = 0; c < 4; ++c) {
      mbar_init(p_full(c), 4);
      mbar_init(p_empty(c), 1);
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

  const int my_items = blockIdx.x < p.items ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto item_geom = [&](int i, int& b, int& h, int& mt) {
    const int idx = blockIdx.x + i * gridDim.x;
    mt = idx % p.mtiles;
    h = (idx / p.mtiles) % p.heads;
    b = idx / (p.mtiles * p.heads);
  };
  const int nck = (p.npad + 63) >> 6;                          // P chunks of 64 keys

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < my_items; ++i) {
        const int buf = i & 1;
        int b, h, mt;
        item_geom(i, b, h, mt);
        mbar_wait(in_empty(buf), ((i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(in_full(buf), Cfg::ITEM);
        const uint32_t base = in_base + buf * Cfg::ITEM;
        tma_load_2d(base, &tmQ, in_full(buf), h * 64, b * p.t + mt * 128);
        tma_load_2d(base + Cfg::Q_BYTES, &tmKV, in_full(buf), (p.heads + h) * 64, b * p.t);
        tma_load_2d(base + Cfg::Q_BYTES + Cfg::KV_BYTES, &tmKV, in_full(buf), (2 * p.heads + h) * 64, b * p.t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, p.npad);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
      auto qk = [&](int i) {                                   // S[i & 1] = Q K^T of item i
        const int buf = i & 1;
        mbar_wait(in_full(buf), (i >> 1) & 1u);
        mbar_wait(s_empty(buf), ((i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t base = in_base + buf * Cfg::ITEM;
        const uint64_t adesc = umma_desc_kmajor(base, 1024u, 2u);
        const uint64_t bdesc = umma_desc_kmajor(base + Cfg::Q_BYTES, 1024u, 2u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + buf * 256, adesc + 2u * k, bdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full(buf));
      };
      if (my_items > 0) qk(0);
      for (int i = 0; i < my_items; ++i) {
        if (i + 1 < my_items) qk(i + 1);
        const int buf = i & 1;
        const uint32_t vbase = in_base + buf * Cfg::ITEM + Cfg::Q_BYTES + Cfg::KV_BYTES;
        const int ksteps = p.npad >> 4;
        for (int c = 0; c < nck; ++c) {
          mbar_wait(p_full(c), i & 1u);                        // chunk c of P(i) written; S(i) columns <= 64c+63 consumed
          tc_fence_after();
          const int k1 = (c * 4 + 4 < ksteps) ? c * 4 + 4 : ksteps;
          for (int ks = c * 4; ks < k1; ++ks) {
            const uint64_t adesc = umma_desc_kmajor(p_base + c * Cfg::P_CHUNK + (ks & 3) * 32, 1024u, 2u);
            const uint64_t bdesc = umma_desc_kmajor(vbase + ks * 2048, 1024u, 2u);    // MN-major: 16 keys = 2 atoms of 8 rows
            umma_bf16(tmem_base + buf * 256, adesc, bdesc, idesc_o, ks != 0 ? 1u : 0u);
          }
          umma_commit(p_empty(c));
        }
        umma_commit(in_empty(buf));
        umma_commit(o_full(buf));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax groups (thread = query row of the group's items)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;                           // items i with i % 2 == grp, S buffer grp
    const int r = q * 32 + lane;
    const uint32_t s_addr = tmem_base + grp * 256 + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunks = (p.npad + 31) >> 5;                    // 32-column TMEM reads per row
    for (int i = grp; i < my_items; i += 2) {
      int b, h, mt;
      item_geom(i, b, h, mt);
      mbar_wait(s_full(grp), (i >> 1) & 1u);
      tc_fence_after();
      // A warp whose 32 rows all lie beyond T (second row tile of a 199-token window: rows 224..255) only keeps the
      // barrier protocol going; its P rows are never read into a stored output row.
      const bool warp_live = mt * 128 + q * 32 < p.t;
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
        const uint32_t chunk = p_base + (c >> 1) * Cfg::P_CHUNK + r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pr[j * 8 + e * 2], pr[j * 8 + e * 2 + 1]);
          st_shared_v4(chunk + ((((c & 1) * 4 + j) ^ (r & 7)) << 4), u);
        }
      };
      auto chunk_done = [&](int c) {                           // after the second half (or the tail) of a 64-key chunk
        if ((c & 1) == 1 || c == nchunks - 1) {
          tc_fence_before();
          fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(c >> 1));
        }
      };
      if (warp_live) tmem_ld_32x32(s_addr, va);
      for (int c = 0; c < nchunks; c += 2) {
        mbar_wait(p_empty(c >> 1), (i & 1u) ^ 1u);            // PV of the previous item is done with this P chunk
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
      mbar_wait(o_full(grp), (i >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(s_addr, v0);
      tmem_ld_32x32(s_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(grp));                // S / O of this buffer may be overwritten by item i + 2
      const int row = mt * 128 + r;
      if (row < p.t) {
        const float inv_l = 1.0f / l;
        __nv_bfloat16* orow = p.out + ((long long)b * p.t + row) * (p.heads * 64) + h * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h2[e] = __floats2bfloat162_rn(__uint_as_float(v0[j * 8 + e * 2]) * inv_l, __uint_as_float(v0[j * 8 + e * 2 + 1]) * inv_l);
          reinterpret_cast<uint4*>(orow)[j] = u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h2[e] = __floats2bfloat162_rn(__uint_as_float(v1[j * 8 + e * 2]) * inv_l, __uint_as_float(v1[j * 8 + e * 2 + 1]) * inv_l);
          reinterpret_cast<uint4*>(orow)[4 + j] = u;
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
