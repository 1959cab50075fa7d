// 3x3 "same" stride-1 convolution on tcgen05 with the input halo staged ONCE in shared memory
// (ResNet-50 bottleneck conv2 of layer1 / layer2: reference src/architectures/video.py:22-28, 3x3 with
// padding="same", folded BN + ReLU).
//
// The generic implicit-GEMM kernel (tc_gemm.cuh) fetches one activation box per filter tap, i.e. every input
// pixel crosses L2 -> shared memory nine times; with 64 / 128 output channels those layers are bound by that
// traffic (measured 12-14 TB/s of L2 reads), not by the tensor pipe.  Here a CTA tile is `bh` whole output rows
// of one image.  One TMA box per 64-channel chunk brings the (bh+2) x (W+2) input window -- the zero padding is
// TMA out-of-bounds fill -- into a 128B-swizzled slab whose rows are the pixels of the PADDED-pitch raster
// (pitch P = W + 2).  In that raster a filter tap (dy, dx) is a pure shift of dy*P + dx rows, so the A operand of
// tap (dy, dx) is the same slab read through a UMMA descriptor whose start address is advanced by that many
// 128-byte rows.  (The 128B swizzle is a function of the absolute shared-memory address, which is also how TMA
// wrote the slab, so the descriptor's base-offset field stays 0 -- measured on B200: setting it to the row phase
// gives wrong results.)  Output row m = ro*P + xo of the accumulator is pixel (ro, xo); the two pad columns per
// row are computed and dropped in the epilogue.
//
//   tile      : 256 accumulator rows (two M=128 MMAs per tap and K step) = bh = floor(256 / P) image rows
//   A traffic : (bh+2)*P rows per 64 channels, once           (was 9 x 128 rows per 128 outputs)
//   B traffic : Cout = 64: all 9*C/64 weight tiles resident in shared memory for the whole kernel;
//               Cout = 128: streamed through a ring, one tile per (chunk, tap) feeding 2 MMAs x 4 K steps
//   epilogue  : tcgen05.ld -> bias + ReLU -> bf16 -> dense [bh][W] 128B-swizzled staging -> one TMA store per
//               64 output channels, clipped at the image bottom (direct 16-byte global stores from the pixel-owning
//               threads measured 8 % slower on layer1 and bought nothing from the extra activation stages they free)
// Warp roles: warp0 activation producer, warp1 MMA issuer, warp2 TMEM allocator, warp3 weight producer, warps 4-11 epilogue
// (warps 4-7 rows 0-127, warps 8-11 rows 128-255).  Persistent, one CTA per SM, accumulators double-buffered.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

struct Conv3Params {
  int H, W, NB, C, Cout;
  int P;            // padded pitch W + 2
  int bh;           // output rows per tile
  int tiles_h;      // ceil(H / bh)
  int num_tiles;    // NB * tiles_h
  int reverse;      // 1: tiles are walked from the last to the first
  int kchunks;      // C / 64
  unsigned a_bytes; // bytes of one activation box: (bh + 2) * P * 128
  unsigned a_stage; // bytes reserved per activation stage: rows read by the MMAs (2P + 2 + 256), rounded up to 1 KB
  int a_stages;     // activation stages (2..4)
  const float* bias;
  int act;
};

template <int BN, bool B_RESIDENT>
struct Conv3Cfg {
  static constexpr int MAX_A_STAGES = 4;
  static constexpr int B_TILE = BN * 128;                    // [Cout rows][64 channels] K-major
  static constexpr int B_SLOTS = B_RESIDENT ? 9 : 4;         // resident: 9 taps x kchunks (kchunks == 1); else a ring
  static constexpr int B_BYTES = B_SLOTS * B_TILE;
  static constexpr int C_CHUNK = 256 * 128;                  // dense staging of one 64-channel chunk (<= 256 pixels)
  static constexpr int C_CHUNKS = BN / 64;
  static constexpr int C_BYTES = C_CHUNKS * C_CHUNK;
  static constexpr int BUDGET = 226 * 1024;                  // dynamic shared memory incl. 1 KB alignment slack
  static constexpr int TMEM_COLS = 4 * BN;                   // 2 accumulator buffers x 2 row halves x BN columns
  static constexpr int THREADS = 384;
  static_assert(BN == 64 || BN == 128, "conv3x3: Cout tile must be 64 or 128");
};

template <int BN, bool B_RESIDENT>
__global__ void __launch_bounds__(384, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Conv3Params p) {
  using Cfg = Conv3Cfg<BN, B_RESIDENT>;
  extern __shared__ uint8_t smem_raw[];
  // barriers: fullA[4] emptyA[4] tfull[2] tempty[2] fullB[B_SLOTS] emptyB[B_SLOTS]
  __shared__ __align__(8) uint64_t bars[12 + 2 * Cfg::B_SLOTS];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base;
  const uint32_t c_base = b_base + Cfg::B_BYTES;
  const uint32_t a_base = c_base + Cfg::C_BYTES;
  const uint32_t bar_base = smem_u32(bars);
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (10 + a); };
  auto fullB = [&](int s) { return bar_base + 8u * (12 + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (12 + Cfg::B_SLOTS + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::MAX_A_STAGES; ++s) {
      mbar_init(fullA(s), 1);
      mbar_init(emptyA(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 8);           // one arrive per epilogue warp
    }
    for (int s = 0; s < Cfg::B_SLOTS; ++s) {
      mbar_init(fullB(s), 1);
      mbar_init(emptyB(s), 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------ activation producer (runs a_stages slabs ahead)
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
        const int n = te / p.tiles_h;
        const int h0 = (te - n * p.tiles_h) * p.bh;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(emptyA(sa), pa ^ 1u);
          mbar_arrive_expect_tx(fullA(sa), p.a_bytes);
          tma_load_5d(a_base + sa * p.a_stage, &tmA, fullA(sa), kc * 64, -1, h0 - 1, n, 0);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ weight producer (own warp: never blocks the slab loads)
    if (lane == 0) {
      if (B_RESIDENT) {                                  // the whole filter bank, once (weights are constants)
        mbar_arrive_expect_tx(fullB(0), 9u * Cfg::B_TILE);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_base + tap * Cfg::B_TILE, &tmB, fullB(0), tap * p.C, 0);
      } else {
        int sb = 0;
        uint32_t pb = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(emptyB(sb), pb ^ 1u);
              mbar_arrive_expect_tx(fullB(sb), Cfg::B_TILE);
              tma_load_2d(b_base + sb * Cfg::B_TILE, &tmB, fullB(sb), tap * p.C + kc * 64, 0);
              if (++sb == Cfg::B_SLOTS) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      if (B_RESIDENT) {
        mbar_wait(fullB(0), 0);
        tc_fence_after();
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        mbar_wait(tempty(acc), ((local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 2 * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(fullA(sa), pa);
          tc_fence_after();
          const uint32_t slab = a_base + sa * p.a_stage;
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t b_addr;
            if (B_RESIDENT) {
              b_addr = b_base + tap * Cfg::B_TILE;
            } else {
              mbar_wait(fullB(sb), pb);
              tc_fence_after();
              b_addr = b_base + sb * Cfg::B_TILE;
            }
            const uint64_t bdesc = umma_desc_kmajor(b_addr, 1024u, 2u);
            const uint32_t row_off = static_cast<uint32_t>((tap / 3) * p.P + (tap % 3)) * 128u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint64_t adesc = umma_desc_kmajor(slab + row_off + half * 128u * 128u, 1024u, 2u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem + half * BN, adesc + 2u * k, bdesc + 2u * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
            }
            if (!B_RESIDENT) {
              umma_commit(emptyB(sb));
              if (++sb == Cfg::B_SLOTS) { sb = 0; pb ^= 1u; }
            }
          }
          umma_commit(emptyA(sa));
          if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
        }
        umma_commit(tfull(acc));
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: accumulator rows -> pixels of bh image rows
    const int half = (warp - 4) >> 2;
    const int q = warp & 3;
    const int m = half * 128 + q * 32 + lane;          // accumulator row = position in the padded-pitch raster
    const int ro = m / p.P, xo = m - ro * p.P;
    const bool valid = (xo < p.W) && (ro < p.bh);
    const int dpix = ro * p.W + xo;                    // dense pixel index inside the tile's [bh][W] output box
    const uint32_t st_row = static_cast<uint32_t>(dpix) * 128u;
    const uint32_t st_xor = static_cast<uint32_t>(dpix & 7);
    const bool store_thread = (threadIdx.x == 128);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
      const int n = te / p.tiles_h;
      const int h0 = (te - n * p.tiles_h) * p.bh;
      mbar_wait(tfull(acc), (local >> 1) & 1u);
      tc_fence_after();
      if (store_thread) bulk_wait_group_read<0>();      // previous tile's stores have read the staging buffers
      named_bar_sync(1, 256);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * 2 * BN + half * BN + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (valid) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + j));
              f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
            }
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] < 0.0f ? 0.0f : f[j];   // NaN-propagating like torch.relu
          }
          const uint32_t cbuf = c_base + (c >> 1) * Cfg::C_CHUNK + st_row;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
            st_shared_v4(cbuf + ((((c & 1) * 4 + j) ^ st_xor) << 4), u);
          }
        }
      }
      // accumulator drained: hand it back before the store handshake
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      fence_proxy_async();
      named_bar_sync(2, 256);
      if (store_thread) {
#pragma unroll
        for (int cc = 0; cc < Cfg::C_CHUNKS; ++cc) tma_store_5d(&tmC, c_base + cc * Cfg::C_CHUNK, cc * 64, 0, h0, n, 0);
        bulk_commit_group();
      }
    }
    if (store_thread) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
