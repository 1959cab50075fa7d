// ResNet-50 stem fused with its max-pool: 7x7/2 "TF-same" conv (pad 2|3) + folded BN + ReLU + 3x3/2 un-padded
// max-pool, [n,232,240,4] zero-bordered bf16 crops -> [n,55,55,64] bf16
// (reference src/architectures/video.py:63-90 Conv2dSame, :98-103, :116-117).
//
// Why a dedicated kernel: run as two launches the 112x112x64 stem output (1.6 MB per crop) is written to HBM and
// read back by the pool (822 MB per 256 crops), and the generic kernel re-fetches the 28 KB filter bank for every
// output row through 7 tiny pipeline stages, which leaves it latency-bound at 190 TFLOP/s.  Here
//   * the filter bank (7 filter rows x 4 KB, UMMA core-matrix order) is loaded ONCE per CTA and stays resident;
//   * one pipeline stage = the 7 padded input rows of one output row (7 x 1920 B, SWIZZLE_NONE strips whose
//     stride-2 window overlap is expressed in the UMMA descriptor: LBO 16 B, SBO 128 B), 8 stages in flight;
//   * a work unit is 14 pooled rows of one crop = 29 consecutive stem rows; the epilogue keeps the last four stem
//     rows (ReLU'd, bf16) in a shared-memory ring and emits pooled row j as soon as stem rows 2j..2j+2 are there,
//     so the stem activation never leaves the SM.
// max-pool of the bf16-rounded ReLU outputs == bf16 rounding of the pooled fp32 values (rounding is monotonic), so
// the result is bit-identical to the two-kernel path; NaNs propagate like torch (ReLU select, __hmax2_nan).
// Warp roles: warp0 strip producer, warp1 MMA issuer, warp2 TMEM allocator, warps 4-11 epilogue + pooling.
#pragma once
#include "tc_gemm.cuh"

namespace avcer {

struct StemPoolParams {
  int n;                    // crops
  int units;                // n * 4 work units (14 pooled rows each; the last of a crop has 13)
  const void* w_packed;     // [7][4 KB] filter rows in UMMA no-swizzle core-matrix order
  const float* bias;        // [64] folded BN shift
  __nv_bfloat16* out;       // [n, 55, 55, 64] with `out_pitch` elements between pixels
  long long out_pitch;      // >= 64, multiple of 8 (a wider row lets the caller place other channels next to the pooled ones)
};

struct StemPoolCfg {
  static constexpr int STRIP = 2048;                 // one padded input row (1920 B) per 2 KB slot
  static constexpr int A_STAGE = 7 * STRIP;          // the 7 filter rows of one output row
  static constexpr int STAGES = 8;
  static constexpr int B_BYTES = 7 * 4096;
  static constexpr int ROW = 112 * 128;              // one stem output row: 112 px x 64 ch bf16
  static constexpr int RING = 4;
  static constexpr int SMEM = STAGES * A_STAGE + 1024 /*junk-row overread of the last strip*/ + B_BYTES + RING * ROW + 1024;
  static constexpr int TMEM_COLS = 128;              // 2 accumulator buffers x 64 columns
  static constexpr int THREADS = 384;
  static constexpr int UNIT_ROWS = 14;               // pooled rows per unit
};

// Epilogue + pooling of the stem kernels (warps 4-11): bias + ReLU -> 4-row shared-memory ring -> 3x3/2 max-pool.
// tbar: shared address of the tfull[2] / tempty[2] barrier block.
__device__ __forceinline__ void stem_pool_epilogue(const StemPoolParams& p, uint32_t tmem_base, uint32_t ring_base, uint32_t tbar,
                                                   int warp, int lane) {
  using Cfg = StemPoolCfg;
  auto tfull_bar = [&](int a) { return tbar + 8u * a; };
  auto tempty_bar = [&](int a) { return tbar + 8u * (2 + a); };
  auto unit_geom = [&](int unit, int& n, int& j0, int& rows) {
    n = unit >> 2;
    j0 = (unit & 3) * Cfg::UNIT_ROWS;
    rows = (j0 + Cfg::UNIT_ROWS <= 55) ? Cfg::UNIT_ROWS : 55 - j0;
  };
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;                 // which 32 of the 64 output channels
    const int px = q * 32 + lane;                    // stem pixel (accumulator row); 112..127 are junk rows
    const int et = threadIdx.x - 128;                // 0..255
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + grp * 32 + j);
    int local = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int n, j0, rows;
      unit_geom(unit, n, j0, rows);
      for (int r = 0; r <= 2 * rows; ++r, ++local) {
        const int acc = local & 1;
        mbar_wait(tfull_bar(acc), (local >> 1) & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * 64 + grp * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));           // accumulator is free again
        if (px < 112) {
          const uint32_t row = ring_base + (local & 3) * Cfg::ROW + px * 128;   // ring position runs on across units
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = __uint_as_float(v[j * 8 + e * 2]) + bias[j * 8 + e * 2];
              float b = __uint_as_float(v[j * 8 + e * 2 + 1]) + bias[j * 8 + e * 2 + 1];
              a = a < 0.0f ? 0.0f : a;                           // NaN-propagating like torch.relu
              b = b < 0.0f ? 0.0f : b;
              h2[e] = __floats2bfloat162_rn(a, b);
            }
            st_shared_v4(row + (((grp * 4 + j) ^ (px & 7)) << 4), u);
          }
        }
        named_bar_sync(1, 256);                                   // stem row r is complete in the ring
        if (r >= 2 && (r & 1) == 0) {
          // pooled row j = j0 + r/2 - 1 from stem rows r-2, r-1, r: 55 pixels x 8 chunks of 8 channels
          const int j = j0 + (r >> 1) - 1;
          __nv_bfloat16* orow = p.out + ((long long)(n * 55 + j) * 55) * p.out_pitch;
          for (int it = et; it < 55 * 8; it += 256) {
            const int po = it >> 3, ch = it & 7;
            uint4 m;
            __nv_bfloat162* mm = reinterpret_cast<__nv_bfloat162*>(&m);
            bool first = true;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t rbase = ring_base + ((local - 2 + dy) & 3) * Cfg::ROW;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int sp = 2 * po + dx;
                uint4 u;
                ld_shared_v4(rbase + sp * 128 + ((ch ^ (sp & 7)) << 4), u);
                const __nv_bfloat162* uu = reinterpret_cast<const __nv_bfloat162*>(&u);
                if (first) {
                  m = u;
                  first = false;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) mm[e] = __hmax2_nan(mm[e], uu[e]);
                }
              }
            }
            *reinterpret_cast<uint4*>(orow + po * p.out_pitch + ch * 8) = m;
          }
        }
      }
    }
}

__global__ void __launch_bounds__(384, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap tmA, const StemPoolParams p) {
  using Cfg = StemPoolCfg;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::STAGES + 5];     // full[8] empty[8] tfull[2] tempty[2] bfull
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + Cfg::STAGES * Cfg::A_STAGE + 1024;
  const uint32_t ring_base = b_base + Cfg::B_BYTES;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);          // one arrive per epilogue warp
    }
    mbar_init(bfull_bar, 1);
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

  // unit -> crop, first pooled row, pooled rows in the unit; stem rows r0 .. r0 + 2*rows (inclusive)
  auto unit_geom = [&](int unit, int& n, int& j0, int& rows) {
    n = unit >> 2;
    j0 = (unit & 3) * Cfg::UNIT_ROWS;
    rows = (j0 + Cfg::UNIT_ROWS <= 55) ? Cfg::UNIT_ROWS : 55 - j0;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ strip producer
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull_bar, Cfg::B_BYTES);           // filter bank: constant, fetched once
      bulk_load_1d(b_base, p.w_packed, Cfg::B_BYTES, bfull_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int n, j0, rows;
        unit_geom(unit, n, j0, rows);
        for (int r = 2 * j0; r <= 2 * (j0 + rows); ++r) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), 7u * 1920u);
          for (int ky = 0; ky < 7; ++ky)
            tma_load_5d(a_base + stage * Cfg::A_STAGE + ky * Cfg::STRIP, &tmA, full_bar(stage), 0, 0, r, n, ky);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: 14 x (128 x 64 x 16) per stem row
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      mbar_wait(bfull_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int n, j0, rows;
        unit_geom(unit, n, j0, rows);
        for (int r = 0; r <= 2 * rows; ++r, ++local) {
          const int acc = local & 1;
          mbar_wait(tempty_bar(acc), ((local >> 1) & 1u) ^ 1u);
          tc_fence_after();
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            // A: output pixel ox reads the 32 elements starting 16 B * ox into the strip (core matrices 8 rows x 16 B,
            // rows 16 B apart: LBO = 16 B to the next K core, SBO = 128 B to the next 8 rows)
            const uint64_t adesc = umma_desc_nosw(a_base + stage * Cfg::A_STAGE + ky * Cfg::STRIP, 16u, 128u);
            // B: packed core matrices [k/8][n/8]: LBO (next k core) = 64/8 * 128 B, SBO (next n core) = 128 B
            const uint64_t bdesc = umma_desc_nosw(b_base + ky * 4096, 64u * 16u, 128u);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 128u * k, idesc, (ky | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(acc));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    stem_pool_epilogue(p, tmem_base, ring_base, bar_base + 8u * (2 * Cfg::STAGES), warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// K1 fused into the stem: the same stem + pool, fed with the uint8 BGR crops themselves ([n,224,224,3], packed 224x224:
// the resize of data/utils.py:19-39 is the identity, BASELINE configs 1-4).  K1's whole job for such crops -- u8 -> fp32,
// subtract the channel means, round to bf16, NHWC4 with a zero border -- happens on the way into shared memory, so the
// 442 KB per crop that K1 writes and the stem re-reads never exist: the stem reads 150 KB of pixels per crop.
//   * shared memory holds a ring of 16 padded input rows (2 KB each, the same SWIZZLE_NONE strip layout K1 produced in
//     HBM; borders zeroed once).  Stem row r multiplies rows 2r .. 2r+6, so advancing one stem row costs two new rows.
//   * one producer thread streams the image rows (672 contiguous bytes each) into a 16-slot raw ring with 1-D bulk
//     copies, up to 16 rows ahead; four converter warps take turns on the row pairs a stem row adds: 28 lanes x 4 pixel
//     pairs per row, exactly K1's arithmetic (float(px) - mean, __floats2bfloat162_rn: bit-identical strips),
//     conflict-free 16-byte stores, one generic->async proxy fence per pass, then one arrive per row `full` barrier.
//   * the MMA thread waits for the two newest rows, issues the same 14 UMMAs per stem row as stem_pool_kernel and
//     commits to the `empty` barriers of the two rows that leave the 7-row window.
// Epilogue (bias, ReLU, 4-row ring, 3x3/2 max-pool) is shared with stem_pool_kernel: outputs are bit-identical.
struct StemPoolU8Cfg {
  static constexpr int STRIP = 2048;
  static constexpr int RING_ROWS = 16;               // power of two: slot = g & 15, use count = g >> 4
  static constexpr int A_BYTES = RING_ROWS * STRIP;
  static constexpr int B_BYTES = 7 * 4096;
  static constexpr int ROW = 112 * 128;
  static constexpr int RING = 4;
  static constexpr int RAW_ROW = 704;                // one image row: 672 bytes of BGR pixels in a 704-byte slot
  static constexpr int RAW_BYTES = RING_ROWS * RAW_ROW;
  static constexpr int SMEM = A_BYTES + 1024 /*junk-row overread of the last slot*/ + B_BYTES + RING * ROW + RAW_BYTES + 1024;
  static constexpr int TMEM_COLS = 128;
  static constexpr int THREADS = 448;                // 12 warps of stem_pool_kernel + 2 more converter warps
  static constexpr int UNIT_ROWS = 14;
  static constexpr int CONV_WARPS = 4;               // warps 2, 3, 12, 13
};

__global__ void __launch_bounds__(448, 1)
stem_pool_u8_kernel(const uint8_t* __restrict__ crops, const StemPoolParams p) {
  using Cfg = StemPoolU8Cfg;
  extern __shared__ uint8_t smem_raw[];
  // full[16] empty[16] tfull[2] tempty[2] bfull rawfull[16] rawempty[16]
  __shared__ __align__(8) uint64_t bars[4 * Cfg::RING_ROWS + 5];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + Cfg::A_BYTES + 1024;
  const uint32_t ring_base = b_base + Cfg::B_BYTES;
  const uint32_t raw_base = ring_base + Cfg::RING * Cfg::ROW;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::RING_ROWS + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::RING_ROWS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::RING_ROWS + 2 + a); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * Cfg::RING_ROWS + 4);
  auto rawfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::RING_ROWS + 5 + s); };
  auto rawempty_bar = [&](int s) { return bar_base + 8u * (3 * Cfg::RING_ROWS + 5 + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::RING_ROWS; ++s) {
      mbar_init(full_bar(s), 1);            // one arrive from the converter warp that owns the row
      mbar_init(empty_bar(s), 1);           // one tcgen05.commit when the row has left the 7-row window
      mbar_init(rawfull_bar(s), 1);         // the bulk copy of the row's 672 bytes (or a plain arrive for a padding row)
      mbar_init(rawempty_bar(s), 1);        // the converter warp has the bytes in registers
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);
    }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_slot_s), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  // zero the strip ring once: the left / right borders of a padded row (pixels 0-1 and 226-239) are never written again
  for (uint32_t i = threadIdx.x; i < (Cfg::A_BYTES + 1024) / 16; i += blockDim.x)
    st_shared_v4(a_base + i * 16, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);
  pdl_wait();
  pdl_launch_dependents();

  auto unit_geom = [&](int unit, int& n, int& j0, int& rows) {
    n = unit >> 2;
    j0 = (unit & 3) * Cfg::UNIT_ROWS;
    rows = (j0 + Cfg::UNIT_ROWS <= 55) ? Cfg::UNIT_ROWS : 55 - j0;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ raw-row producer: one 672-byte bulk copy per image row,
    // up to 16 rows ahead of the converters (global latency is hidden here, not in the converter warps)
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull_bar, Cfg::B_BYTES);           // filter bank: constant, fetched once
      bulk_load_1d(b_base, p.w_packed, Cfg::B_BYTES, bfull_bar);
      uint32_t g = 0;                                           // ring-row counter (32-bit unsigned: slot = g & 15, use = g >> 4)
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int n, j0, rows;
        unit_geom(unit, n, j0, rows);
        const int p0 = 4 * j0, np = 4 * rows + 7;               // padded rows 4*j0 .. 4*(j0+rows)+6 (stem rows 2*j0 ..)
        const uint8_t* img = crops + (size_t)n * (224 * 224 * 3);
        for (int i = 0; i < np; ++i, ++g) {
          const int slot = (int)(g & (Cfg::RING_ROWS - 1));
          mbar_wait(rawempty_bar(slot), ((g >> 4) & 1u) ^ 1u);
          const int y = p0 + i - 2;                             // image row of padded row p0 + i
          if (y >= 0 && y < 224) {
            mbar_arrive_expect_tx(rawfull_bar(slot), 672u);
            bulk_load_1d(raw_base + slot * Cfg::RAW_ROW, img + (size_t)y * 672, 672u, rawfull_bar(slot));
          } else {
            mbar_arrive(rawfull_bar(slot));                     // TF-"same" padding row: nothing to fetch
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 12) {
    // ------------------------------------------------------------ converters: four warps, one PASS each in turn.
    // A pass = the two ring rows a stem row adds to the window (the first seven rows of a unit go as 2 + 2 + 2 + 1):
    // 28 lanes x 4 pixel pairs per row, conflict-free 16-byte stores, ONE generic->async proxy fence per pass (the fence
    // is the expensive part: per row and per lane it left the first versions at 215-240 us against 164 + 33 us).
    const int cw = warp < 4 ? warp - 2 : warp - 10;             // 0..3
    const float m0 = 91.4953f, m1 = 103.8827f, m2 = 131.0912f;  // data/utils.py:27-29 (B, G, R)
    uint32_t g0 = 0;                                            // ring row of the unit's first padded row
    uint32_t pass = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int n, j0, rows;
      unit_geom(unit, n, j0, rows);
      const int p0 = 4 * j0;
      const int npass = 4 + 2 * rows;
      for (int k = 0; k < npass; ++k, ++pass) {
        if ((int)(pass & (Cfg::CONV_WARPS - 1)) != cw) continue;
        const int i0 = k < 4 ? 2 * k : 7 + 2 * (k - 4);         // first row of the pass inside the unit
        const int nr = k == 3 ? 1 : 2;
        for (int rr = 0; rr < nr; ++rr) {
          const uint32_t g = g0 + i0 + rr;
          const int slot = (int)(g & (Cfg::RING_ROWS - 1));
          const uint32_t use = g >> 4;
          const int y = p0 + i0 + rr - 2;                       // image row of padded row p0 + i
          const bool real = y >= 0 && y < 224;
          mbar_wait(rawfull_bar(slot), use & 1u);
          uint32_t px16[4][3];                                  // 4 pixel pairs x 6 bytes
          if (real && lane < 28) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t src = raw_base + slot * Cfg::RAW_ROW + (q * 28 + lane) * 6;
#pragma unroll
              for (int h = 0; h < 3; ++h) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(px16[q][h]) : "r"(src + 2 * h));
            }
            fence_proxy_async();                                // generic-proxy reads ordered before the bulk-copy refill
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(rawempty_bar(slot));       // the raw slot may be refilled
          mbar_wait(empty_bar(slot), (use & 1u) ^ 1u);
          const uint32_t dst = a_base + slot * Cfg::STRIP + 2 * 8;   // pixel 2 of the strip (8 bytes per NHWC4 pixel)
          if (lane < 28) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {                       // two pixels per 16-byte store
              uint4 u = make_uint4(0u, 0u, 0u, 0u);
              if (real) {
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
                const uint32_t a = px16[q][0], b = px16[q][1], c = px16[q][2];
                h2[0] = __floats2bfloat162_rn((float)(a & 0xffu) - m0, (float)(a >> 8) - m1);
                h2[1] = __floats2bfloat162_rn((float)(b & 0xffu) - m2, 0.f);
                h2[2] = __floats2bfloat162_rn((float)(b >> 8) - m0, (float)(c & 0xffu) - m1);
                h2[3] = __floats2bfloat162_rn((float)(c >> 8) - m2, 0.f);
              }
              st_shared_v4(dst + (q * 28 + lane) * 16, u);
            }
          }
        }
        fence_proxy_async();                                    // the pass's strip stores -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0)
          for (int rr = 0; rr < nr; ++rr) mbar_arrive(full_bar((int)((g0 + i0 + rr) & (Cfg::RING_ROWS - 1))));
      }
      g0 += 4 * rows + 7;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: 14 x (128 x 64 x 16) per stem row
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      mbar_wait(bfull_bar, 0);
      tc_fence_after();
      uint32_t g0 = 0;                                           // ring-row counter of the unit's first padded row (32-bit unsigned)
      int local = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int n, j0, rows;
        unit_geom(unit, n, j0, rows);
        const int nrows = 2 * rows + 1;                          // stem rows of the unit
        for (int r = 0; r < nrows; ++r, ++local) {
          const int acc = local & 1;
          mbar_wait(tempty_bar(acc), ((local >> 1) & 1u) ^ 1u);
          tc_fence_after();
          // rows g0+2r .. g0+2r+6; all but the two newest were waited for by the previous stem row
          for (int ky = (r == 0 ? 0 : 5); ky < 7; ++ky) {
            const uint32_t g = g0 + 2 * r + ky;
            mbar_wait(full_bar((int)(g & (Cfg::RING_ROWS - 1))), (g >> 4) & 1u);
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            const uint32_t g = g0 + 2 * r + ky;
            const uint64_t adesc = umma_desc_nosw(a_base + (g & (Cfg::RING_ROWS - 1)) * Cfg::STRIP, 16u, 128u);
            const uint64_t bdesc = umma_desc_nosw(b_base + ky * 4096, 64u * 16u, 128u);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 128u * k, idesc, (ky | k) != 0 ? 1u : 0u);
          }
          // rows that leave the window: two per stem row, all seven after the unit's last stem row
          const int nfree = (r == nrows - 1) ? 7 : 2;
          for (int f = 0; f < nfree; ++f) umma_commit(empty_bar((int)((g0 + 2 * r + f) & (Cfg::RING_ROWS - 1))));
          umma_commit(tfull_bar(acc));
        }
        g0 += 4 * rows + 7;
      }
    }
  } else if (warp >= 4) {
    stem_pool_epilogue(p, tmem_base, ring_base, bar_base + 8u * (2 * Cfg::RING_ROWS), warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace avcer
