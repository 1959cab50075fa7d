// Baseline-JPEG decode of the face crops on the GPU (image half): replaces the
// `cv2.imread` at src/get_prob_video.py:95, i.e. libjpeg-turbo's decompressor with the defaults OpenCV leaves in place
// (JDCT_ISLOW, fancy up-sampling), for the files src/data/get_face_images.py:60 writes with `cv2.imwrite` defaults:
// baseline sequential DCT, 8 bit, YCbCr 4:2:0 (4:4:4 also handled), no restart markers.  Output is bit-identical to
// cv2.imread: every stage is the library's integer arithmetic, restated (oracle/jpeg.py pins the same restatement against
// cv2.imdecode on the CPU).
//
// Three kernels, all integer / byte work bound by latency and HBM, not by math:
//   0. jpeg_unstuff_kernel  -- the byte stuffing of the entropy-coded segments (FF 00 -> FF) removed by a block-per-image
//      stream compaction (the host only walks the marker segments: ~20 per file);
//   1. jpeg_huffman_kernel  -- ONE THREAD PER IMAGE walks its entropy-coded segment (T.81 F.2.2): 64-bit bit buffer refilled
//      with aligned 32-bit loads, 10-bit look-ahead tables in shared memory (code length + symbol in one lookup, canonical
//      max-code search for the few longer codes), one symbol per loop iteration so that the 32 images of a warp stay
//      in lock-step; non-zero coefficients are scattered into a zeroed int16 [block][64] array in natural order.
//      A 224x224 crop at quality 95 is ~35 000 symbols; the batch supplies the parallelism (thousands of crops per call).
//   2. jpeg_idct_kernel     -- jidctint.c jpeg_idct_islow: 8 threads per block, pass 1 on columns (de-quantising on the way
//      in), 8x8 transpose through shared memory, pass 2 on rows, post-IDCT range limit (mod-1024 table semantics),
//      8-byte row stores into per-component sample planes.
//   3. jpeg_color_kernel    -- jdsample.c h2v2_fancy_upsample (triangle filter with the +8 / +7 rounding alternation, first /
//      last chroma row and column handled like the library's context rows) + jdcolor.c ycc_rgb_convert (16-bit fixed-point
//      tables as arithmetic, range limiting), written as packed BGR -- the byte layout cv2.imread returns and K1 consumes.
#include "common.h"

namespace avcer {

struct JpegImage {            // mirrors avcer_jpeg_image (include/avcer_b200.h)
  long long data_off;         // byte offset of the un-stuffed segment in `data` (multiple of 4); the stuffed one starts at raw + data_off + src_shift
  long long data_len;         // its length in bytes
  long long coef_off;         // first block of the image in the coefficient array (Y blocks, then Cb, then Cr)
  long long plane_off;        // byte offset of the Y plane in `planes` (Cb, Cr follow)
  long long out_off;          // byte offset of the BGR image in `out`
  int width, height;
  int mcus_w, mcus_h;
  int hs;                     // luma sampling factor per axis: 2 = 4:2:0, 1 = 4:4:4
  int qt_y, qt_c;             // quantisation table indices
  int src_shift;              // 0..3: the segment may sit at any byte of `raw` (whole files uploaded as they are)
};

constexpr int LOOK = 10;      // look-ahead bits of the fast Huffman path

struct HuffTab {              // one Huffman table expanded in shared memory
  unsigned short look[1 << LOOK];    // (code length << 8) | symbol, 0 = longer than LOOK bits
  int maxcode[18];                   // canonical decoding (T.81 F.2.2.3): largest code of each length, -1 = none
  int valoff[17];                    // vals index of the first code of each length minus that code
  unsigned char vals[256];
};

// natural (row-major) index of zig-zag position k
__constant__ unsigned char c_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                                          28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                                          54, 47, 55, 62, 63};
__device__ __forceinline__ int zigzag_natural(int k) { return c_zigzag[k]; }

// Byte un-stuffing (T.81 B.1.1.5: a 0x00 follows every 0xFF data byte) as a per-image stream compaction: one block per
// image, 1 KB chunks, warp-shuffle + shared-memory exclusive scan of the keep flags, scatter.  A restart marker
// (FF D0..D7) inside the segment is reported through `status` (not covered: cv2.imwrite never writes them by default).
constexpr int UNSTUFF_THREADS = 256;
__global__ void __launch_bounds__(UNSTUFF_THREADS)
jpeg_unstuff_kernel(const unsigned char* __restrict__ raw, const JpegImage* __restrict__ imgs, unsigned char* __restrict__ data,
                    long long* __restrict__ lens, int* __restrict__ status) {
  __shared__ int warp_sums[UNSTUFF_THREADS / 32];
  __shared__ long long base_s;
  const int img = blockIdx.x;
  const long long off = imgs[img].data_off, len = imgs[img].data_len;
  const unsigned char* src = raw + off + imgs[img].src_shift;
  unsigned char* dst = data + off;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (long long c0 = 0; c0 < len; c0 += UNSTUFF_THREADS * 4) {
    // every thread owns 4 consecutive bytes
    const long long i0 = c0 + threadIdx.x * 4;
    unsigned char b[4];
    int keep[4], cnt = 0;
    unsigned char prev = (i0 > 0 && i0 - 1 < len) ? src[i0 - 1] : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + j;
      b[j] = i < len ? src[i] : 0;
      keep[j] = i < len && !(prev == 0xFF && b[j] == 0x00);
      if (i < len && prev == 0xFF && b[j] >= 0xD0 && b[j] <= 0xD7) atomicExch(status, -(img + 1));     // restart marker
      // a stuffed 0x00 is dropped, and does not itself make the NEXT byte look stuffed (FF 00 00 keeps the second 00)
      prev = b[j];
      cnt += keep[j];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < UNSTUFF_THREADS / 32; ++w) {
      if (w < wid) wbase += warp_sums[w];
      total += warp_sums[w];
    }
    long long pos = base_s + wbase + incl - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (keep[j]) dst[pos++] = b[j];
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    lens[img] = base_s;
    for (long long i = base_s; i < ((base_s + 3) & ~3ll); ++i) dst[i] = 0;   // the bit reader loads whole words (stays inside this image's slot)
  }
}

// bits: [4][16] code counts per length, vals: [4][256] symbols; table order DC0, AC0, DC1, AC1
constexpr int HUFF_THREADS = 32;      // one warp per block: a batch of 1 500 crops already spreads over 47 SMs

__global__ void __launch_bounds__(HUFF_THREADS)
jpeg_huffman_kernel(const unsigned char* __restrict__ data, const JpegImage* __restrict__ imgs, int n,
                    const long long* __restrict__ lens, const unsigned char* __restrict__ bits,
                    const unsigned char* __restrict__ vals, short* __restrict__ coefs, int* __restrict__ status, int ipw) {
  __shared__ HuffTab tabs[4];
  // ---- expand the four tables (every thread block builds its own copy; ~4 K entries)
  for (int t = 0; t < 4; ++t) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tabs[t].vals[i] = vals[t * 256 + i];
    for (int i = threadIdx.x; i < (1 << LOOK); i += blockDim.x) tabs[t].look[i] = 0;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    HuffTab& T = tabs[threadIdx.x];
    const unsigned char* b = bits + threadIdx.x * 16;
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
      const int cnt = b[len - 1];
      T.valoff[len] = k - code;
      for (int c = 0; c < cnt; ++c, ++code, ++k) {
        if (len <= LOOK) {
          const int lo = code << (LOOK - len);
          for (int f = 0; f < (1 << (LOOK - len)); ++f) T.look[lo + f] = (unsigned short)((len << 8) | T.vals[k]);
        }
      }
      T.maxcode[len] = cnt ? code - 1 : -1;
      code <<= 1;
    }
    T.maxcode[17] = 0x7fffffff;
  }
  __syncthreads();

  // `ipw` images per warp (lanes >= ipw idle): a small batch is spread over more warps, so that every SM has work, the
  // lanes of a warp diverge less (a block costs the maximum symbol count over the warp's images) and fewer lanes' cache
  // misses serialise on one warp.  ipw = 32 once the batch is large enough to fill the machine anyway.
  if ((int)threadIdx.x >= ipw) return;
  const int img = blockIdx.x * ipw + threadIdx.x;
  if (img >= n) return;
  const JpegImage im = imgs[img];
  const unsigned int* words = reinterpret_cast<const unsigned int*>(data + im.data_off);
  const long long n_words = (lens[img] + 3) >> 2;             // un-stuffed length (jpeg_unstuff_kernel)
  long long wi = 0;
  unsigned long long bb = 0;      // bit buffer, MSB-aligned content in the low `nb` bits
  int nb = 0;
  unsigned int ahead = n_words > 0 ? __ldg(words) : 0u;       // the next word is always already on its way
  auto refill = [&]() {           // keep at least 32 valid bits (zero bits past the end, like libjpeg's padding)
    if (nb <= 32) {
      unsigned int w = ahead;
      ++wi;
      ahead = wi < n_words ? __ldg(words + wi) : 0u;
      w = __byte_perm(w, 0, 0x0123);                       // big-endian bit order
      bb = (bb << 32) | w;
      nb += 32;
    }
  };
  auto peek = [&](int k) -> unsigned int { return (unsigned int)(bb >> (nb - k)) & ((1u << k) - 1u); };

  const int blocks_per_mcu = im.hs * im.hs + 2;
  const int yw = im.mcus_w * im.hs;                         // luma blocks per row
  const long long y_blocks = (long long)yw * im.mcus_h * im.hs;
  const long long c_blocks = (long long)im.mcus_w * im.mcus_h;
  const long long n_mcus = c_blocks;
  int pred_y = 0, pred_cb = 0, pred_cr = 0;                   // DC predictors in registers (no dynamically indexed array)
  int bad = 0;
  for (long long mcu = 0; mcu < n_mcus && !bad; ++mcu) {
    const int mx = (int)(mcu % im.mcus_w), my = (int)(mcu / im.mcus_w);
    for (int b = 0; b < blocks_per_mcu && !bad; ++b) {
      int comp;
      long long blk;
      if (b < im.hs * im.hs) {
        comp = 0;
        blk = (long long)(my * im.hs + b / im.hs) * yw + mx * im.hs + b % im.hs;
      } else {
        comp = b - im.hs * im.hs + 1;
        blk = y_blocks + (comp - 1) * c_blocks + (long long)my * im.mcus_w + mx;
      }
      short* out = coefs + (im.coef_off + blk) * 64;
      const HuffTab& dc = tabs[comp == 0 ? 0 : 2];
      const HuffTab& ac = tabs[comp == 0 ? 1 : 3];
      int k = 0;
      // one symbol per iteration: DC (k == 0) then AC run/size pairs until EOB or k == 64
      while (k < 64) {
        refill();
        const HuffTab& T = k == 0 ? dc : ac;
        int len, sym;
        const unsigned int e = T.look[peek(LOOK)];
        if (e) {
          len = e >> 8;
          sym = e & 255;
        } else {                                            // codes longer than LOOK bits: canonical search
          len = LOOK + 1;
          int code = (int)peek(len);
          while (len <= 16 && code > T.maxcode[len]) { ++len; code = (int)peek(len); }
          if (len > 16) { bad = 1; break; }
          sym = T.vals[(T.valoff[len] + code) & 255];
        }
        nb -= len;
        const int size = sym & 15, run = sym >> 4;
        int v = 0;
        if (size) {
          refill();
          v = (int)peek(size);
          nb -= size;
          if (v < (1 << (size - 1))) v -= (1 << size) - 1;  // EXTEND (T.81 F.2.2.1)
        }
        if (k == 0) {
          int& pr = comp == 0 ? pred_y : comp == 1 ? pred_cb : pred_cr;
          pr += v;
          out[0] = (short)pr;
          k = 1;
        } else if (size == 0) {
          if (run != 15) break;                             // EOB
          k += 16;                                          // ZRL
        } else {
          k += run;
          if (k > 63) { bad = 1; break; }
          out[zigzag_natural(k)] = (short)v;
          ++k;
        }
      }
    }
  }
  if (bad) atomicExch(status, img + 1);
}

// ------------------------------------------------------------------ jidctint.c jpeg_idct_islow
constexpr int CONST_BITS = 13, PASS1_BITS = 2;
__device__ __forceinline__ void idct8(const int (&d)[8], int (&o)[8], int shift) {
  int z2 = d[2], z3 = d[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = d[0]; z3 = d[4];
  int tmp0 = (z2 + z3) << CONST_BITS;
  int tmp1 = (z2 - z3) << CONST_BITS;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = d[7]; tmp1 = d[5]; tmp2 = d[3]; tmp3 = d[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  o[0] = (tmp10 + tmp3 + rnd) >> shift; o[7] = (tmp10 - tmp3 + rnd) >> shift;
  o[1] = (tmp11 + tmp2 + rnd) >> shift; o[6] = (tmp11 - tmp2 + rnd) >> shift;
  o[2] = (tmp12 + tmp1 + rnd) >> shift; o[5] = (tmp12 - tmp1 + rnd) >> shift;
  o[3] = (tmp13 + tmp0 + rnd) >> shift; o[4] = (tmp13 - tmp0 + rnd) >> shift;
}
// sample_range_limit + CENTERJSAMPLE indexed by (x & RANGE_MASK) (jdmaster.c prepare_range_limit_table)
__device__ __forceinline__ unsigned int range_limit_idct(int x) {
  const int idx = x & 1023;
  return idx < 128 ? idx + 128 : idx < 512 ? 255 : idx < 896 ? 0 : idx - 896;
}

// block_img: for every 256-block chunk the image it starts in is found by binary search over coef_off
__global__ void __launch_bounds__(256)
jpeg_idct_kernel(const short* __restrict__ coefs, const JpegImage* __restrict__ imgs, int n, long long total_blocks,
                 const unsigned short* __restrict__ qtabs, unsigned char* __restrict__ planes) {
  __shared__ int ws[32][8][9];
  const int lb = threadIdx.x >> 3, c = threadIdx.x & 7;       // local block, column (pass 1) / row (pass 2)
  const long long blk = (long long)blockIdx.x * 32 + lb;
  const bool live = blk < total_blocks;
  int lo = 0;
  if (live) {
    int hi = n - 1;                                            // last image whose coef_off <= blk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (imgs[mid].coef_off <= blk) lo = mid; else hi = mid - 1;
    }
  }
  const JpegImage im = imgs[live ? lo : 0];
  const long long rel = blk - im.coef_off;
  const int yw = im.mcus_w * im.hs;
  const long long y_blocks = (long long)yw * im.mcus_h * im.hs, c_blocks = (long long)im.mcus_w * im.mcus_h;
  int comp = 0, bw = yw;
  long long r2 = rel;
  if (rel >= y_blocks) { comp = 1 + (int)((rel - y_blocks) / c_blocks); r2 = (rel - y_blocks) % c_blocks; bw = im.mcus_w; }
  if (live) {
    const unsigned short* q = qtabs + (comp == 0 ? im.qt_y : im.qt_c) * 64;
    const short* cf = coefs + blk * 64;
    int d[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) d[r] = (int)cf[r * 8 + c] * (int)q[r * 8 + c];
    idct8(d, o, CONST_BITS - PASS1_BITS);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[lb][r][c] = o[r];
  }
  __syncthreads();
  if (live) {
    int d[8], o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = ws[lb][c][j];
    idct8(d, o, CONST_BITS + PASS1_BITS + 3);
    const long long y_bytes = (long long)yw * 8 * im.mcus_h * im.hs * 8, c_bytes = (long long)im.mcus_w * 8 * im.mcus_h * 8;
    unsigned char* plane = planes + im.plane_off + (comp == 0 ? 0 : y_bytes + (comp - 1) * c_bytes);
    const int by = (int)(r2 / bw), bx = (int)(r2 % bw);
    uint2 u;
    u.x = range_limit_idct(o[0]) | (range_limit_idct(o[1]) << 8) | (range_limit_idct(o[2]) << 16) | (range_limit_idct(o[3]) << 24);
    u.y = range_limit_idct(o[4]) | (range_limit_idct(o[5]) << 8) | (range_limit_idct(o[6]) << 16) | (range_limit_idct(o[7]) << 24);
    *reinterpret_cast<uint2*>(plane + ((long long)(by * 8 + c) * bw + bx) * 8) = u;
  }
}

// ------------------------------------------------------------------ jdsample.c h2v2_fancy_upsample + jdcolor.c ycc_rgb_convert
__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

__global__ void __launch_bounds__(256)
jpeg_color_kernel(const JpegImage* __restrict__ imgs, int n, const long long* __restrict__ pix_prefix,
                  const unsigned char* __restrict__ planes, unsigned char* __restrict__ out, long long total_pixels) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_pixels) return;
  int lo = 0, hi = n - 1;                                      // last image whose first pixel <= p
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (pix_prefix[mid] <= p) lo = mid; else hi = mid - 1;
  }
  const JpegImage im = imgs[lo];
  const long long rel = p - pix_prefix[lo];
  const int y = (int)(rel / im.width), x = (int)(rel % im.width);
  const int yw = im.mcus_w * im.hs * 8, cw = im.mcus_w * 8;    // plane pitches
  const long long y_bytes = (long long)yw * im.mcus_h * im.hs * 8, c_bytes = (long long)cw * im.mcus_h * 8;
  const unsigned char* Y = planes + im.plane_off;
  const unsigned char* Cb = Y + y_bytes;
  const unsigned char* Cr = Cb + c_bytes;
  const int yy = Y[(long long)y * yw + x];
  int cb, cr;
  if (im.hs == 1) {
    cb = Cb[(long long)y * cw + x];
    cr = Cr[(long long)y * cw + x];
  } else {
    const int rows = (im.height + 1) >> 1, cols = (im.width + 1) >> 1;
    const int cy = y >> 1, cx = x >> 1;
    const int oy = (y & 1) ? min(cy + 1, rows - 1) : max(cy - 1, 0);    // the farther chroma row (itself at the borders)
    auto up = [&](const unsigned char* C) -> int {
      const unsigned char* r0 = C + (long long)cy * cw;
      const unsigned char* r1 = C + (long long)oy * cw;
      const int s = 3 * r0[cx] + r1[cx];
      if ((x & 1) == 0) {
        if (cx == 0) return (s * 4 + 8) >> 4;
        return (3 * s + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
      }
      if (cx == cols - 1) return (s * 4 + 7) >> 4;
      return (3 * s + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
    };
    cb = up(Cb);
    cr = up(Cr);
  }
  // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768
  const int xb = cb - 128, xr = cr - 128;
  const int r = clamp255(yy + ((91881 * xr + 32768) >> 16));
  const int b = clamp255(yy + ((116130 * xb + 32768) >> 16));
  const int g = clamp255(yy + ((-22554 * xb + 32768 - 46802 * xr) >> 16));
  unsigned char* o = out + im.out_off + ((long long)y * im.width + x) * 3;
  o[0] = (unsigned char)b; o[1] = (unsigned char)g; o[2] = (unsigned char)r;
}

}  // namespace avcer

using namespace avcer;

static_assert(sizeof(JpegImage) == sizeof(avcer_jpeg_image), "JpegImage must mirror avcer_jpeg_image");

extern "C" int avcer_jpeg_decode(const uint8_t* raw, const avcer_jpeg_image* images, int n, const uint8_t* huff_bits,
                                 const uint8_t* huff_vals, const uint16_t* qtables, const int64_t* pixel_prefix,
                                 int64_t total_blocks, int64_t total_pixels, uint8_t* data, int64_t* lens, int16_t* coefs,
                                 uint8_t* planes, uint8_t* out, int32_t* status, void* stream) {
  AVCER_REQUIRE(n >= 0 && total_blocks >= 0 && total_pixels >= 0, "jpeg_decode: negative size");
  if (n == 0) return 0;
  AVCER_REQUIRE(raw && data && lens && images && huff_bits && huff_vals && qtables && pixel_prefix && coefs && planes && out && status,
                "jpeg_decode: null pointer");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(data) & 3) == 0 && (reinterpret_cast<uintptr_t>(planes) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(coefs) & 1) == 0,
                "jpeg_decode: data must be 4-byte, planes 8-byte aligned");
  AVCER_REQUIRE(total_blocks < (1ll << 36), "jpeg_decode: too many blocks");
  cudaStream_t st = as_stream(stream);
  AVCER_CUDA(cudaMemsetAsync(coefs, 0, (size_t)total_blocks * 64 * sizeof(int16_t), st));
  AVCER_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  const JpegImage* imgs = reinterpret_cast<const JpegImage*>(images);
  jpeg_unstuff_kernel<<<n, UNSTUFF_THREADS, 0, st>>>(raw, imgs, data, reinterpret_cast<long long*>(lens), status);
  if (int rc = check_launch("jpeg_unstuff_kernel")) return rc;
  int ipw = (n + 4 * num_sms() - 1) / (4 * num_sms());      // aim at >= 4 warps per SM
  ipw = ipw < 1 ? 1 : ipw > HUFF_THREADS ? HUFF_THREADS : ipw;
  jpeg_huffman_kernel<<<(n + ipw - 1) / ipw, HUFF_THREADS, 0, st>>>(data, imgs, n, reinterpret_cast<const long long*>(lens), huff_bits,
                                                                    huff_vals, coefs, status, ipw);
  if (int rc = check_launch("jpeg_huffman_kernel")) return rc;
  jpeg_idct_kernel<<<(unsigned)((total_blocks + 31) / 32), 256, 0, st>>>(coefs, imgs, n, total_blocks, qtables, planes);
  if (int rc = check_launch("jpeg_idct_kernel")) return rc;
  jpeg_color_kernel<<<(unsigned)((total_pixels + 255) / 256), 256, 0, st>>>(imgs, n, reinterpret_cast<const long long*>(pixel_prefix), planes,
                                                                             out, total_pixels);
  return check_launch("jpeg_color_kernel");
}
