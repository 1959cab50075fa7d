// Face-detector (RetinaFace-ResNet50) kernels that the VS kernels do not already cover -- SURVEY.md section 8(f) row 4,
// reference: src/data/face_detection/ibug/face_detection/retina_face/{retina_face_predictor.py, retina_face_net.py,
// box_utils.py, prior_box.py} behind src/data/get_face_images.py:38-63.  The 1x1 / 3x3 convolutions of the body, FPN, SSH
// and heads run on avcer_contract (tcgen05 implicit GEMM); here are
//   * the stem on raw uint8 video frames: mean subtraction + conv 7x7/2 pad 3 (BatchNorm folded) + ReLU,
//   * max-pool 3x3/2 pad 1 (torchvision ResNet-50),
//   * FPN top-down merge: a + nearest-upsampled b,
//   * anchors + 2-class softmax + box / landmark decoding of the three head maps into the predictor's [P, 15] rows.
#include "common.h"
#include "ptx.cuh"
#include "vec8.cuh"

namespace avcer {

// ------------------------------------------------------------------ stem: uint8 BGR/RGB frame -> [n, H/2, W/2, 64]
// (retina_face_predictor.py:61-67: image.astype(int) - (104, 117, 123) in BGR order, then torchvision resnet50 conv1 + bn1 +
// relu.)  Direct convolution in fp32: a block computes 8 x 32 output pixels x 64 channels from a 21 x 69 pixel patch held in
// shared memory as mean-subtracted floats (zero outside the frame: the padding applies AFTER the subtraction) and the
// 147 x 64 folded filter bank; a thread owns one pixel and all 64 channels, so every filter value it reads (a 16-byte
// broadcast) feeds 4 FMAs.
constexpr int ST_TW = 32, ST_TH = 8, ST_PW = 2 * (ST_TW - 1) + 7, ST_PH = 2 * (ST_TH - 1) + 7;   // 69 x 21 patch
constexpr int ST_PATCH = (ST_PH * ST_PW * 3 + 3) / 4 * 4;       // floats, rounded so that the filter bank behind it is 16-byte aligned
constexpr int ST_SMEM = (ST_PATCH + 147 * 64) * 4;

template <typename T>
__global__ void __launch_bounds__(256)
det_stem_kernel(const uint8_t* __restrict__ frames, int h, int w, int rgb, const float* __restrict__ wt,
                const float* __restrict__ bias, T* __restrict__ out, int ho, int wo) {
  extern __shared__ __align__(16) float st_smem[];
  float* patch = st_smem;                               // [21][69][3]
  float* sw = st_smem + ST_PATCH;                       // [147][64]
  for (int i = threadIdx.x; i < 147 * 64; i += 256) sw[i] = wt[i];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.z, oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  const uint8_t* src = frames + (long long)b * h * w * 3;
  const float mean[3] = {104.f, 117.f, 123.f};
  for (int i = threadIdx.x; i < ST_PH * ST_PW * 3; i += 256) {
    const int c = i % 3, px = (i / 3) % ST_PW, py = i / (3 * ST_PW);
    const int y = 2 * oy0 - 3 + py, x = 2 * ox0 - 3 + px;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) v = (float)src[((long long)y * w + x) * 3 + (rgb ? 2 - c : c)] - mean[c];
    patch[i] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int oy = oy0 + ty, ox = ox0 + tx;
  float acc[64];
#pragma unroll
  for (int co = 0; co < 64; ++co) acc[co] = __ldg(bias + co);
  const float* prow = patch + ((2 * ty) * ST_PW + 2 * tx) * 3;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
#pragma unroll 3
    for (int t = 0; t < 21; ++t) {                      // (kx, c) of this filter row: 21 consecutive patch floats
      const float v = prow[ky * ST_PW * 3 + t];
      const float4* wr = reinterpret_cast<const float4*>(sw + (ky * 21 + t) * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 w4 = wr[q];
        acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
      }
    }
  }
  if (oy < ho && ox < wo) {
    T* dst = out + (((long long)b * ho + oy) * wo + ox) * 64;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(acc[8 * q + e], 0.f);
      Vec8<T>::store(dst + 8 * q, v);
    }
  }
}

// ------------------------------------------------------------------ stem input for the tensor-core path
// uint8 frame -> zero-bordered NHWC4 [n, hp, wp, 4] in the 16-bit storage type: input pixel (y, x) lands at (y + 3, x + 3)
// as (B - 104, G - 117, R - 123, 0) -- integers of magnitude <= 151, exact in bf16 and fp16 -- everything else is zero, so
// that the 7x7 / 2 pad-3 stem becomes the strip-mode contraction the VS stem uses (avcer_contract a_strip: output ox of a
// filter row reads the 8 pixels x 4 channels starting at padded column 2 ox; pixel 7 and channel 3 carry zero weights).
__global__ void det_prepare_kernel(const uint8_t* __restrict__ frames, int n, int h, int w, int rgb, int hp, int wp,
                                   __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * hp * wp) return;
  const int px = (int)(i % wp), py = (int)((i / wp) % hp), b = (int)(i / ((long long)wp * hp));
  const int x = px - 3, y = py - 3;
  float v[3] = {0.f, 0.f, 0.f};
  if (x >= 0 && x < w && y >= 0 && y < h) {
    const uint8_t* s = frames + (((long long)b * h + y) * w + x) * 3;
    v[0] = (float)s[rgb ? 2 : 0] - 104.f;
    v[1] = (float)s[1] - 117.f;
    v[2] = (float)s[rgb ? 0 : 2] - 123.f;
  }
  uint2 u;
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
  h2[0] = __floats2bfloat162_rn(v[0], v[1]);
  h2[1] = __floats2bfloat162_rn(v[2], 0.f);
  reinterpret_cast<uint2*>(out)[i] = u;
}

// ------------------------------------------------------------------ max pool 3x3/2 pad 1 (torchvision resnet50.maxpool)
template <typename T>
__global__ void maxpool3x3s2p1_kernel(const T* __restrict__ x, int n, int h, int w, int c, int ho, int wo, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = c / 8;
  if (i >= (long long)n * ho * wo * c8) return;
  const int cc = (int)(i % c8) * 8;
  const int ox = (int)((i / c8) % wo);
  const int oy = (int)((i / ((long long)c8 * wo)) % ho);
  const int b = (int)(i / ((long long)c8 * wo * ho));
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = 2 * oy - 1 + dy;
    if (yy < 0 || yy >= h) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = 2 * ox - 1 + dx;
      if (xx < 0 || xx >= w) continue;
      float v[8];
      Vec8<T>::load(x + (((long long)b * h + yy) * w + xx) * c + cc, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = (v[j] > m[j] || v[j] != v[j]) ? v[j] : m[j];
    }
  }
  Vec8<T>::store(y + (((long long)b * ho + oy) * wo + ox) * c + cc, m);
}

// ------------------------------------------------------------------ FPN merge: out = a + nearest_upsample(b)
// (retina_face_net.py:88-94: F.interpolate(mode="nearest") to a's size; the source row / column of every output row /
// column comes from the host, computed with PyTorch's own index rule.)
template <typename T>
__global__ void upsample_add_kernel(const T* __restrict__ a, const T* __restrict__ bsrc, int n, int h, int w, int hb, int wb, int c,
                                    const int* __restrict__ ymap, const int* __restrict__ xmap, T* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = c / 8;
  if (i >= (long long)n * h * w * c8) return;
  const int cc = (int)(i % c8) * 8;
  const int x = (int)((i / c8) % w);
  const int y = (int)((i / ((long long)c8 * w)) % h);
  const int b = (int)(i / ((long long)c8 * w * h));
  float va[8], vb[8];
  Vec8<T>::load(a + (((long long)b * h + y) * w + x) * c + cc, va);
  Vec8<T>::load(bsrc + (((long long)b * hb + ymap[y]) * wb + xmap[x]) * c + cc, vb);
#pragma unroll
  for (int j = 0; j < 8; ++j) va[j] += vb[j];
  Vec8<T>::store(out + (((long long)b * h + y) * w + x) * c + cc, va);
}

// ------------------------------------------------------------------ anchors + softmax + decode -> [n, P, 15]
// One thread per prior.  heads_k: [n * fh_k * fw_k, pitch] fp32 rows of one pyramid level, columns
// [cls a0 (bg, face), cls a1 | box a0 (4), box a1 | landmarks a0 (10), a1].  Row layout of the result = the predictor's:
// x1, y1, x2, y2 (pixels), score, 5 x (x, y).  Every fp32 operation is spelled out in the reference's order
// (prior_box.py:24-29 in double then rounded; box_utils.py:223-227, 243-248; retina_face_predictor.py:75-84) without
// FMA contraction; only expf differs from torch's by an ulp or two.
struct DecodeParams {
  const float* heads[3];
  long long pitch;
  int n, height, width;
  int fh[3], fw[3];
  long long P;
  float* dets;
};

__global__ void det_decode_kernel(const DecodeParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n * p.P) return;
  const int b = (int)(i / p.P);
  long long q = i % p.P;
  int k = 0;
  for (; k < 2; ++k) {
    const long long cnt = 2ll * p.fh[k] * p.fw[k];
    if (q < cnt) break;
    q -= cnt;
  }
  const int a = (int)(q & 1);
  const long long cell = q >> 1;
  const int col = (int)(cell % p.fw[k]), row = (int)(cell / p.fw[k]);
  const int step = 8 << k;
  const double min_size = (double)((16 << (2 * k)) << a);                  // 16, 32 | 64, 128 | 256, 512
  const float pcx = (float)((col + 0.5) * step / p.width), pcy = (float)((row + 0.5) * step / p.height);
  const float pw = (float)(min_size / p.width), ph = (float)(min_size / p.height);
  const float* hrow = p.heads[k] + ((long long)b * p.fh[k] * p.fw[k] + cell) * p.pitch;
  const float l0 = hrow[2 * a], l1 = hrow[2 * a + 1];
  const float m = fmaxf(l0, l1);
  const float e0 = expf(__fsub_rn(l0, m)), e1 = expf(__fsub_rn(l1, m));
  const float score = __fdiv_rn(e1, __fadd_rn(e0, e1));
  const float* loc = hrow + 4 + 4 * a;
  const float* lmk = hrow + 12 + 10 * a;
  const float W = (float)p.width, H = (float)p.height;
  float* o = p.dets + i * 15;
  const float cx = __fadd_rn(pcx, __fmul_rn(__fmul_rn(loc[0], 0.1f), pw));
  const float cy = __fadd_rn(pcy, __fmul_rn(__fmul_rn(loc[1], 0.1f), ph));
  const float bw = __fmul_rn(pw, expf(__fmul_rn(loc[2], 0.2f)));
  const float bh = __fmul_rn(ph, expf(__fmul_rn(loc[3], 0.2f)));
  const float x1 = __fsub_rn(cx, __fmul_rn(bw, 0.5f)), y1 = __fsub_rn(cy, __fmul_rn(bh, 0.5f));
  o[0] = __fmul_rn(x1, W);
  o[1] = __fmul_rn(y1, H);
  o[2] = __fmul_rn(__fadd_rn(bw, x1), W);
  o[3] = __fmul_rn(__fadd_rn(bh, y1), H);
  o[4] = score;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    o[5 + 2 * j] = __fmul_rn(__fadd_rn(pcx, __fmul_rn(__fmul_rn(lmk[2 * j], 0.1f), pw)), W);
    o[6 + 2 * j] = __fmul_rn(__fadd_rn(pcy, __fmul_rn(__fmul_rn(lmk[2 * j + 1], 0.1f), ph)), H);
  }
}

}  // namespace avcer

using namespace avcer;
typedef __nv_bfloat16 bf16;

#define AVCER_DISPATCH(dtype, ...)                                              \
  do {                                                                          \
    if ((dtype) == AVCER_BF16) { using T = bf16; __VA_ARGS__; }                 \
    else if ((dtype) == AVCER_F32) { using T = float; __VA_ARGS__; }            \
    else return set_error("unknown dtype %d", (int)(dtype));                    \
  } while (0)

extern "C" int avcer_det_stem(const uint8_t* frames, int n, int h, int w, int rgb, const float* wt, const float* bias, void* out,
                              int dtype, void* stream) {
  AVCER_REQUIRE(n >= 0 && h >= 1 && w >= 1, "det_stem: bad shape n=%d h=%d w=%d", n, h, w);
  AVCER_REQUIRE(frames != nullptr && wt != nullptr && bias != nullptr && out != nullptr, "det_stem: null pointer");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0, "det_stem: wt / out must be 16-byte aligned");
  if (n == 0) return 0;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const dim3 grid((wo + ST_TW - 1) / ST_TW, (ho + ST_TH - 1) / ST_TH, n);
  AVCER_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "det_stem: frame too tall or batch too large");
  static bool attr_done[2] = {false, false};
  AVCER_DISPATCH(dtype, {
    auto kern = det_stem_kernel<T>;
    bool& done = attr_done[sizeof(T) == 2 ? 0 : 1];
    if (!done) {
      AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
      done = true;
    }
    launch_pdl(kern, grid, 256, ST_SMEM, as_stream(stream), frames, h, w, rgb, wt, bias, (T*)out, ho, wo);
  });
  return check_launch("det_stem");
}

extern "C" int avcer_det_prepare(const uint8_t* frames, int n, int h, int w, int rgb, int hp, int wp, void* out, void* stream) {
  AVCER_REQUIRE(n >= 0 && h >= 1 && w >= 1 && hp >= h + 6 && wp >= w + 6, "det_prepare: bad shape");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, "det_prepare: out must be 8-byte aligned");
  const long long total = (long long)n * hp * wp;
  if (total == 0) return 0;
  launch_pdl(det_prepare_kernel, blocks_for(total, 256), 256, 0, as_stream(stream), frames, n, h, w, rgb, hp, wp, (__nv_bfloat16*)out);
  return check_launch("det_prepare");
}

extern "C" int avcer_maxpool3x3s2p1(const void* x, int n, int h, int w, int c, void* y, int dtype, void* stream) {
  AVCER_REQUIRE(c % 8 == 0 && h >= 1 && w >= 1, "maxpool3x3s2p1: bad shape");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(maxpool3x3s2p1_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), (const T*)x, n, h, w, c,
                                    ho, wo, (T*)y)));
  return check_launch("maxpool3x3s2p1");
}

extern "C" int avcer_upsample_add(const void* a, const void* b, int n, int h, int w, int hb, int wb, int c, const int32_t* ymap,
                                  const int32_t* xmap, void* out, int dtype, void* stream) {
  AVCER_REQUIRE(c % 8 == 0 && h >= 1 && w >= 1 && hb >= 1 && wb >= 1, "upsample_add: bad shape");
  const long long total = (long long)n * h * w * (c / 8);
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(upsample_add_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), (const T*)a, (const T*)b, n,
                                    h, w, hb, wb, c, (const int*)ymap, (const int*)xmap, (T*)out)));
  return check_launch("upsample_add");
}

extern "C" int avcer_det_decode(const float* heads0, const float* heads1, const float* heads2, int64_t head_pitch, int n, int height,
                                int width, float* dets, void* stream) {
  AVCER_REQUIRE(n >= 0 && height >= 1 && width >= 1 && head_pitch >= 32, "det_decode: bad arguments");
  DecodeParams p{};
  p.heads[0] = heads0; p.heads[1] = heads1; p.heads[2] = heads2;
  p.pitch = head_pitch; p.n = n; p.height = height; p.width = width; p.dets = dets;
  p.P = 0;
  for (int k = 0; k < 3; ++k) {
    const int step = 8 << k;
    p.fh[k] = (height + step - 1) / step;
    p.fw[k] = (width + step - 1) / step;
    p.P += 2ll * p.fh[k] * p.fw[k];
  }
  const long long total = p.n * p.P;
  if (total == 0) return 0;
  launch_pdl(det_decode_kernel, blocks_for(total, 256), 256, 0, as_stream(stream), p);
  return check_launch("det_decode");
}
