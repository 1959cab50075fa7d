// wav2vec2 feature-extractor layer 0 on the tensor cores: Conv1d(1 -> 512, k = 10, stride 5, bias) + LayerNorm(512) + GELU
// (HF Wav2Vec2LayerNormConvLayer #0; called through architectures/audio_8_cl.py:135,180), [B, 64000] fp32 -> [B, 12799, 512].
//
// Why: the SIMT kernel (layers.cu w2v_conv0_kernel) spends ~40 issue slots per output value -- the 10-tap FMAs, their
// shared-memory weight loads, LayerNorm and the erf-GELU -- and is issue-bound at 0.2 of the HBM roof.  Here the
// convolution itself costs nothing: it is a 128 x 512 x 32 contraction per tile,
//     A[t, :] = [x_hi(10) | x_lo(10) | x_hi(10) | 1 | 1]       (the 10 samples of time step t, split in two 16-bit halves)
//     B[c, :] = [w_hi(10) | w_hi(10) | w_lo(10) | b_hi | b_lo]  (filter c and its bias, split the same way)
// i.e. a bf16x3 product (x_hi w_hi + x_lo w_hi + x_hi w_lo + bias, fp32 accumulation: ~16 mantissa bits, the fp32 kernel's
// accuracy class) issued as four tcgen05.mma 128 x 256 x 16.  The A tile is built by two producer warps directly in the UMMA
// no-swizzle core-matrix layout ([k/8][t/8][8 rows][8 elements]: 64 bytes per row as four 16-byte stores; the stride-5
// overlap of consecutive windows is 10 bytes, which no TMA box or descriptor can express); the packed filter bank (32 KB)
// is resident.  The whole 512-channel row of a time step sits in TMEM (128 lanes x 512 columns), so LayerNorm needs no
// second pass over memory: eight epilogue warps (thread = row, two warps per lane quarter, 256 columns each) read the
// accumulator once for (sum, sum of squares), exchange two partials per row through shared memory, and read it again to
// normalise, apply GELU, round and stage 64-column slabs for their own TMA stores (clipped at the window's last step).
// TMEM is fully used by one tile, so the MMAs of tile i+1 wait for the epilogue of tile i: ~5 % of a tile's time.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

// IEEE half has 5 exponent bits: the low half of a quiet sample (|x| ~ 1e-3 -> x_lo ~ 2^-21) would be a subnormal with a few
// bits left, and LayerNorm then amplifies the quiet passage to unit variance.  The half build therefore stores 2^11 x_lo and
// the matching filter block 2^-11 w_hi (weights.pack_conv0_tc); bf16 has fp32's exponent range and needs no scaling.
#ifdef AVCER_HALF
constexpr float kLoScale = 2048.0f;
#else
constexpr float kLoScale = 1.0f;
#endif

struct Conv0Params {
  const float* x;           // [n, t_in] fp32
  const void* w_packed;     // [4 k-cores][64 n-cores][8][8] 16-bit: the B operand described above
  const float* gamma;       // [512] LayerNorm weight
  const float* beta;        // [512] LayerNorm bias
  int n, t_in, t_out;
  int tiles_per_row;        // ceil(t_out / 128)
  int num_tiles;
  float eps;
};

// G = epilogue warps per TMEM lane quarter (each owns 512 / G channels of its 32 rows): 2 -> 8 epilogue warps, 4 -> 16.
template <int G>
struct Conv0Cfg {
  static constexpr int A_BYTES = 128 * 64;            // 128 rows x 32 elements
  static constexpr int B_BYTES = 512 * 64;
  static constexpr int SLAB = 32 * 128;               // 32 rows x 64 columns, 128B-swizzled
  static constexpr int EPI_WARPS = 4 * G;
  static constexpr int C_BYTES = EPI_WARPS * 2 * SLAB;
  static constexpr int PART_BYTES = 2 * G * 128 * 8;  // float2 row partials, double-buffered by tile parity
  static constexpr int SMEM = 2 * A_BYTES + B_BYTES + C_BYTES + 2 * 2048 /*gamma, beta*/ + PART_BYTES + 1024;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;   // warp 0: bank loader, 1: MMA, 2-3: A builders, 4..: epilogue
  static constexpr int TMEM_COLS = 512;
  static constexpr int CW = 512 / G;                  // channels per epilogue warp
};

template <int G>
__global__ void __launch_bounds__(Conv0Cfg<G>::THREADS, 1)
w2v_conv0_tc_kernel(const __grid_constant__ CUtensorMap tmY, const Conv0Params p) {
  using Cfg = Conv0Cfg<G>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];           // afull[2] aempty[2] tfull tempty bfull
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + 2 * Cfg::A_BYTES;
  const uint32_t c_base = b_base + Cfg::B_BYTES;
  const uint32_t g_base = c_base + Cfg::C_BYTES;      // gamma [512] f32, then beta [512] f32
  const uint32_t part_base = g_base + 4096;           // float2 [2 tiles in flight][2 halves][128 rows]
  const uint32_t bar_base = smem_u32(bars);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (2 + s); };
  const uint32_t tfull = bar_base + 8u * 4, tempty = bar_base + 8u * 5, bfull = bar_base + 8u * 6;
  float* g_s = reinterpret_cast<float*>(smem_raw + (g_base - smem_u32(smem_raw)));
  float2* part_s = reinterpret_cast<float2*>(smem_raw + (part_base - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmY);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull(s), 2);               // one arrive per builder warp
      mbar_init(aempty(s), 1);              // tcgen05.commit
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, Cfg::EPI_WARPS);      // one arrive per epilogue warp
    mbar_init(bfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_slot_s), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    g_s[i] = p.gamma[i];
    g_s[512 + i] = p.beta[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);
  pdl_wait();
  pdl_launch_dependents();

  const int my_tiles = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    if (lane == 0) {                        // the packed filter bank: constant, fetched once
      mbar_arrive_expect_tx(bfull, Cfg::B_BYTES);
      bulk_load_1d(b_base, p.w_packed, Cfg::B_BYTES, bfull);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: 4 x (128 x 256 x 16) per tile
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
      mbar_wait(bfull, 0);
      tc_fence_after();
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i & 1;
        mbar_wait(afull(s), (i >> 1) & 1u);
        mbar_wait(tempty, (i & 1u) ^ 1u);   // the epilogue has drained the previous tile (TMEM is single-buffered)
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {       // two halves of the 512 channels
#pragma unroll
          for (int k = 0; k < 2; ++k) {     // K = 32 = two steps of 16 (two 8-element cores each)
            // A: [k/8][t/8][8][8]: next K core 16 x 128 B = 2048 B, next 8 rows 128 B
            const uint64_t adesc = umma_desc_nosw(a_base + s * Cfg::A_BYTES + k * 4096, 2048u, 128u);
            // B: [k/8][c/8][8][8]: next K core 64 x 128 B = 8192 B, next 8 channels 128 B; half h starts 32 cores in
            const uint64_t bdesc = umma_desc_nosw(b_base + h * 4096 + k * 16384, 8192u, 128u);
            umma_bf16(tmem_base + h * 256, adesc, bdesc, idesc, k);
          }
        }
        umma_commit(aempty(s));
        umma_commit(tfull);
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------------------ A builders: thread -> rows r and r + 64 of the tile
    const int r0 = (warp - 2) * 32 + lane;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int b = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * 128;
      const int s = i & 1;
      mbar_wait(aempty(s), ((i >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = r0 + rr * 64, t = t0 + r;
        float xv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) xv[k] = t < p.t_out ? __ldg(p.x + (long long)b * p.t_in + 5 * t + k) : 0.f;
        __nv_bfloat16 e[32];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(xv[k]);
          e[k] = hi;
          e[10 + k] = __float2bfloat16_rn((xv[k] - __bfloat162float(hi)) * kLoScale);
          e[20 + k] = hi;
        }
        e[30] = __float2bfloat16_rn(1.0f);
        e[31] = __float2bfloat16_rn(1.0f);
        const uint32_t row = a_base + s * Cfg::A_BYTES + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) st_shared_v4(row + kc * 2048, *reinterpret_cast<const uint4*>(&e[kc * 8]));
      }
      fence_proxy_async();                  // generic-proxy stores -> visible to the UMMA reads
      __syncwarp();
      if (lane == 0) mbar_arrive(afull(s));
    }
  } else {
    // ------------------------------------------------------------ epilogue: LayerNorm + GELU straight from TMEM
    const int q = warp & 3, g = (warp - 4) >> 2;      // lane quarter (rows 32q ..), channel group (CW g ..)
    const int ew = warp - 4;
    const int r = q * 32 + lane;                      // row of the tile owned by this thread
    const uint32_t cslab = c_base + ew * 2 * Cfg::SLAB;
    const uint32_t row_off = lane * 128;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * Cfg::CW;
    int j = 0;                                        // slabs issued so far by this warp (staging ring of 2)
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int b = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * 128;
      mbar_wait(tfull, i & 1u);
      tc_fence_after();
      // pass 1: (sum, sum of squares) of this thread's CW channels
      uint64_t s1p = 0ull, s2p = 0ull;                // packed (even, odd) column partial sums
#pragma unroll 1
      for (int c = 0; c < Cfg::CW / 64; ++c) {        // two TMEM loads in flight per wait: the pass is latency-bound otherwise
        uint32_t v[2][32];
        tmem_ld_32x32(taddr + c * 64, v[0]);
        tmem_ld_32x32(taddr + c * 64 + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const uint64_t f = pack_f32x2(__uint_as_float(v[hh][k]), __uint_as_float(v[hh][k + 1]));
            s1p = fma_f32x2(f, pack_f32x2(1.0f, 1.0f), s1p);
            s2p = fma_f32x2(f, f, s2p);
          }
      }
      float s1, s2, t1, t2;
      unpack_f32x2(s1p, s1, t1);
      unpack_f32x2(s2p, s2, t2);
      s1 += t1;
      s2 += t2;
      // double-buffered by tile parity: a warp that runs ahead writes tile i+1's partials into the other buffer, and cannot
      // reach tile i+2 before its partner has passed the barrier of tile i+1, i.e. has read tile i's
      float2* part = part_s + (i & 1) * (G * 128);
      part[g * 128 + r] = make_float2(s1, s2);
      named_bar_sync(1, 32 * Cfg::EPI_WARPS);         // every channel group of every row has its partials
      s1 = 0.f;
      s2 = 0.f;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) {                // same order in every group: all of them get the same statistics
        const float2 o = part[gg * 128 + r];
        s1 += o.x;
        s2 += o.y;
      }
      const float mean = s1 * (1.0f / 512.0f);
      const float rstd = rsqrtf(fmaxf(s2 * (1.0f / 512.0f) - mean * mean, 0.f) + p.eps);
      const uint64_t rstd2 = pack_f32x2(rstd, rstd), nm2 = pack_f32x2(-mean * rstd, -mean * rstd);
      // pass 2: normalise, GELU, round, stage 64-column slabs, TMA store.  Steps of 32 columns; the TMEM load of the next
      // step is in flight while the current one is computed (its ~300-cycle latency would otherwise be exposed per step).
      constexpr int STEPS = Cfg::CW / 32;
      uint32_t v[2][32];
      tmem_ld_32x32(taddr, v[0]);
#pragma unroll
      for (int st = 0; st < STEPS; ++st) {
        const int cur = st & 1;
        const uint32_t cbuf = cslab + (j & 1) * Cfg::SLAB;
        tmem_ld_wait();                               // v[cur] has landed
        if (st + 1 < STEPS) tmem_ld_32x32(taddr + (st + 1) * 32, v[cur ^ 1]);
        if ((st & 1) == 0) {
          if (lane == 0) bulk_wait_group_read<1>();   // the store that last read this slab is done with it
          __syncwarp();
        }
        const int c0 = g * Cfg::CW + st * 32;
        uint64_t f[16];                               // 16 channel pairs, all independent: the scheduler interleaves them
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = c0 + 2 * e;
          const uint64_t u = fma_f32x2(pack_f32x2(__uint_as_float(v[cur][2 * e]), __uint_as_float(v[cur][2 * e + 1])), rstd2, nm2);
          f[e] = fma_f32x2(u, *reinterpret_cast<const uint64_t*>(&g_s[c]), *reinterpret_cast<const uint64_t*>(&g_s[512 + c]));
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = gelu_erf_fast2(f[e]);
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float lo, hi;
            unpack_f32x2(f[k8 * 4 + e], lo, hi);
            h2[e] = __floats2bfloat162_rn(lo, hi);
          }
          st_shared_v4(cbuf + row_off + ((((st & 1) * 4 + k8) ^ (lane & 7)) << 4), u);
        }
        if (st & 1) {                                 // a 64-column slab is complete
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&tmY, cbuf, c0 - 32, t0 + q * 32, b, 0, 0);    // box 64 channels x 32 steps, clipped at t_out
            bulk_commit_group();
          }
          ++j;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);             // the accumulator may be overwritten
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
