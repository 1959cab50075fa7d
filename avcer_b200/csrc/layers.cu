// Memory-bound and small layers of the VS / VD / A networks, channels-last, templated on the
// storage type (bf16 in the default mode, fp32 in "fp32 mode").  Arithmetic is always fp32.
// Reference call sites are cited per kernel.
#include "common.h"
#include <stdlib.h>
#include "ptx.cuh"
#include "vec8.cuh"

namespace avcer {

constexpr int ACT_GELU_L = AVCER_ACT_GELU;

template <typename T> __device__ __forceinline__ float gelu_for(float x) { return sizeof(T) == 2 ? gelu_erf_fast(x) : gelu_erf(x); }
// GELU of a pair: packed fp32x2 evaluation where the result is stored as bf16, libdevice erff in fp32 mode
template <typename T> __device__ __forceinline__ void gelu_pair(float& a, float& b) {
  if (sizeof(T) == 2) {
    gelu_erf_fast2(a, b);
  } else {
    a = gelu_erf(a);
    b = gelu_erf(b);
  }
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ max pool 3x3/2, no padding (video.py:103)
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, int n, int h, int w, int c, int ho, int wo, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = c / 8;
  const long long total = (long long)n * ho * wo * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8) * 8;
  const int ox = (int)((i / c8) % wo);
  const int oy = (int)((i / ((long long)c8 * wo)) % ho);
  const int b = (int)(i / ((long long)c8 * wo * ho));
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      float v[8];
      Vec8<T>::load(x + (((long long)b * h + (2 * oy + dy)) * w + (2 * ox + dx)) * c + cc, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = (v[j] > m[j] || v[j] != v[j]) ? v[j] : m[j];   // NaN-propagating like torch
    }
  Vec8<T>::store(y + (((long long)b * ho + oy) * wo + ox) * c + cc, m);
}

// ------------------------------------------------------------------ global average pool (video.py:124)
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, int n, int hw, int c, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = c / 8;
  if (i >= (long long)n * c8) return;
  const int cc = (int)(i % c8) * 8;
  const int b = (int)(i / c8);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = 0; p < hw; ++p) {
    float v[8];
    Vec8<T>::load(x + ((long long)b * hw + p) * c + cc, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j];
  }
  const float inv = 1.0f / (float)hw;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] *= inv;
  Vec8<T>::store(y + (long long)b * c + cc, s);
}

// ------------------------------------------------------------------ tiny Linear (+softmax): warp per row
template <typename T>
__global__ void small_linear_kernel(const T* __restrict__ x, long long n, int k, long long ldx, const float* __restrict__ w,
                                    const float* __restrict__ b, int m, int softmax, float* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < k; i += 32) {
    const float xv = to_f32(x[row * ldx + i]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < m) acc[j] = fmaf(xv, w[(long long)j * k + i], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    float mx = -INFINITY;
    for (int j = 0; j < m; ++j) { acc[j] += b ? b[j] : 0.f; mx = fmaxf(mx, acc[j]); }
    if (softmax) {
      float s = 0.f;
      for (int j = 0; j < m; ++j) { acc[j] = expf(acc[j] - mx); s += acc[j]; }
      for (int j = 0; j < m; ++j) acc[j] /= s;
    }
    for (int j = 0; j < m; ++j) y[row * m + j] = acc[j];
  }
}

// ------------------------------------------------------------------ LSTM cell (PyTorch gate order i,f,g,o; video.py:169-185)
template <typename T>
__global__ void lstm_cell_kernel(const float* __restrict__ xproj, const int* __restrict__ xidx,
                                 const float* __restrict__ hproj, float* __restrict__ c, T* __restrict__ h_out,
                                 long long ldh, long long n, int hidden, int first, int split, float* __restrict__ h_f32) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hidden) return;
  const long long r = i / hidden;
  const int j = (int)(i % hidden);
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (xproj) {
    const long long xr = xidx ? xidx[r] : r;
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = xproj[xr * 4 * hidden + q * hidden + j];
  }
  if (hproj) {
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] += hproj[r * 4 * hidden + q * hidden + j];
  }
  const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanhf(g[2]), og = sigmoid_f(g[3]);
  const float cprev = first ? 0.f : c[i];
  const float cn = fg * cprev + ig * gg;
  c[i] = cn;
  const float hv = og * tanhf(cn);
  if (h_f32) h_f32[i] = hv;
  const T hi = from_f32<T>(hv);
  h_out[r * ldh + j] = hi;
  if (split) {
    // bf16x3 operand of the next recurrent GEMM: [hi | lo | hi] against weights packed [w_hi | w_hi | w_lo], i.e.
    // hi*w_hi + lo*w_hi + hi*w_lo -- 16 mantissa bits on both operands, fp32 accumulation in TMEM
    h_out[r * ldh + hidden + j] = from_f32<T>(hv - to_f32(hi));
    h_out[r * ldh + 2 * hidden + j] = hi;
  }
}

// ------------------------------------------------------------------ GRU cell (PyTorch gate order r,z,n; ExprModelV1, audio_8_cl.py:23-29,63)
// xg: rows [*, 3H] fp32 = W_ih x_t + b_ih (row r of this step at xg + (x_row0 + r * x_row_stride) * 3H);
// hg: [n, 3H] fp32 = W_hh h_{t-1} + b_hh.   n_t = tanh(xg_n + r * hg_n);  h_t = (1 - z) * n_t + z * h_{t-1}.
template <typename T>
__global__ void gru_cell_kernel(const float* __restrict__ xg, long long x_row0, long long x_row_stride,
                                const float* __restrict__ hg, float* __restrict__ h_state, T* __restrict__ h_out, long long ldh,
                                int split, T* __restrict__ y, long long y_row0, long long y_row_stride, long long n, int hidden) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * hidden) return;
  const long long r = i / hidden;
  const int j = (int)(i % hidden);
  const float* xr = xg + (x_row0 + r * x_row_stride) * 3 * hidden;
  const float* hr = hg + r * 3 * hidden;
  const float rg = sigmoid_f(xr[j] + hr[j]);
  const float zg = sigmoid_f(xr[hidden + j] + hr[hidden + j]);
  const float ng = tanhf(xr[2 * hidden + j] + rg * hr[2 * hidden + j]);
  const float hv = (1.0f - zg) * ng + zg * h_state[i];
  h_state[i] = hv;
  const T hi = from_f32<T>(hv);
  h_out[r * ldh + j] = hi;
  if (split) {                                     // bf16x3 operand of the next recurrent GEMM (see lstm_cell_kernel)
    h_out[r * ldh + hidden + j] = from_f32<T>(hv - to_f32(hi));
    h_out[r * ldh + 2 * hidden + j] = hi;
  }
  if (y) y[(y_row0 + r * y_row_stride) * hidden + j] = hi;
}

// x fp32 [rows, k] -> bf16 [rows, 3k] = [hi | lo | hi] (bf16x3 split operand, see lstm_cell_kernel)
__global__ void split_bf16x3_kernel(const float* __restrict__ x, long long rows, int k, long long ldx,
                                    __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * k) return;
  const long long r = i / k;
  const int j = (int)(i % k);
  const float v = x[r * ldx + j];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  out[r * ldo + j] = hi;
  out[r * ldo + k + j] = __float2bfloat16_rn(v - __bfloat162float(hi));
  out[r * ldo + 2 * k + j] = hi;
}

// ------------------------------------------------------------------ K5a audio window gather + pad + zero-mean/unit-var
// (get_prob_audio_8_cl.py:78-90, data/utils.py:63-89, HF zero_mean_unit_var_norm: (x-mean)/sqrt(var+1e-7))
__global__ void __launch_bounds__(512)
audio_normalize_kernel(const float* __restrict__ wav, const long long* __restrict__ starts,
                       const long long* __restrict__ ends, int win, int pad_mode, float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ double red[16];
  __shared__ double bc[2];
  const int wi = blockIdx.x;
  const long long s0 = starts[wi];
  long long len = ends[wi] - s0;
  if (len > win) len = win;
  if (len < 0) len = 0;
  const float* src = wav + s0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  auto block_sum = [&](double v) -> double {
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i]; bc[0] = t; }
    __syncthreads();
    return bc[0];
  };
  // pad value
  float padv = 0.f;
  double csum = 0;
  for (long long i = tid; i < len; i += blockDim.x) csum += (double)src[i];
  csum = block_sum(csum);
  if (pad_mode == 0) padv = len > 0 ? (float)(csum / (double)len) : __int_as_float(0x7fc00000);
  auto value = [&](long long i) -> float {
    if (i < len) return src[i];
    if (pad_mode == 2) return len > 0 ? src[i % len] : 0.f;
    return padv;
  };
  double tsum = 0;
  for (long long i = tid; i < win; i += blockDim.x) tsum += (double)value(i);
  tsum = block_sum(tsum);
  const double mean = tsum / (double)win;
  double vs = 0;
  for (long long i = tid; i < win; i += blockDim.x) { const double d = (double)value(i) - mean; vs += d * d; }
  vs = block_sum(vs);
  const float meanf = (float)mean;
  const float inv = 1.0f / sqrtf((float)(vs / (double)win) + 1e-7f);
  float* o = out + (long long)wi * win;
  for (long long i = tid; i < win; i += blockDim.x) o[i] = (value(i) - meanf) * inv;
}

// ------------------------------------------------------------------ K5b wav2vec2 conv0 (Cin=1,k=10,s=5,bias) + LayerNorm(512) + GELU
// (HF Wav2Vec2LayerNormConvLayer #0).  Persistent blocks (filter + LN parameters staged in shared memory once
// per block); one warp per PAIR of consecutive output time steps, so every filter vector read from shared
// memory feeds two steps; lane owns channels {128*q + 4*lane + e}.
template <typename T>
__global__ void __launch_bounds__(256, 3)
w2v_conv0_kernel(const float* __restrict__ x, int n, int t_in, int t_out, const float* __restrict__ w,
                 const float* __restrict__ b, const float* __restrict__ g, const float* __restrict__ be,
                 T* __restrict__ y, long long y_pitch_rows) {
  __shared__ __align__(16) float sw[10][512];
  __shared__ __align__(16) float sb[512], sg[512], sbe[512];
  for (int i = threadIdx.x; i < 5120; i += blockDim.x) sw[i >> 9][i & 511] = w[(i & 511) * 10 + (i >> 9)];   // w is [512][1][10]
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sb[i] = b[i]; sg[i] = g[i]; sbe[i] = be[i]; }
  __syncthreads();
  pdl_wait();                       // the filter / LN parameters above are constants; x and y belong to other kernels
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pairs_per_row = (t_out + 1) >> 1;
  const long long pairs = (long long)n * pairs_per_row;
  for (long long pi = (long long)blockIdx.x * 8 + wid; pi < pairs; pi += (long long)gridDim.x * 8) {
    const int batch = (int)(pi / pairs_per_row);
    const int t = 2 * (int)(pi - (long long)batch * pairs_per_row);
    const bool two = t + 1 < t_out;
    const float* xb = x + (long long)batch * t_in + 5 * t;
    float xv[15];
#pragma unroll
    for (int k = 0; k < 10; ++k) xv[k] = __ldg(xb + k);
#pragma unroll
    for (int k = 10; k < 15; ++k) xv[k] = two ? __ldg(xb + k) : 0.0f;
    float v[2][16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c0 = 128 * q + 4 * lane;
      const float4 bias = *reinterpret_cast<const float4*>(&sb[c0]);
      // packed fp32x2 FMAs (two channels per instruction; per-lane arithmetic identical to fmaf)
      uint64_t a0l = pack_f32x2(bias.x, bias.y), a0h = pack_f32x2(bias.z, bias.w), a1l = a0l, a1h = a0h;
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const float4 wk = *reinterpret_cast<const float4*>(&sw[k][c0]);
        const uint64_t wl = pack_f32x2(wk.x, wk.y), wh = pack_f32x2(wk.z, wk.w);
        const uint64_t x0 = pack_f32x2(xv[k], xv[k]), x1 = pack_f32x2(xv[k + 5], xv[k + 5]);
        a0l = fma_f32x2(wl, x0, a0l); a0h = fma_f32x2(wh, x0, a0h);
        a1l = fma_f32x2(wl, x1, a1l); a1h = fma_f32x2(wh, x1, a1h);
      }
      unpack_f32x2(a0l, v[0][4 * q], v[0][4 * q + 1]); unpack_f32x2(a0h, v[0][4 * q + 2], v[0][4 * q + 3]);
      unpack_f32x2(a1l, v[1][4 * q], v[1][4 * q + 1]); unpack_f32x2(a1h, v[1][4 * q + 2], v[1][4 * q + 3]);
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s0 += v[0][i]; s1 += v[1][i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    const float mean[2] = {s0 * (1.0f / 512.0f), s1 * (1.0f / 512.0f)};
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d0 = v[0][i] - mean[0], d1 = v[1][i] - mean[1];
      q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
    const float rstd[2] = {rsqrtf(q0 * (1.0f / 512.0f) + 1e-5f), rsqrtf(q1 * (1.0f / 512.0f) + 1e-5f)};
    T* yr = y + ((long long)batch * y_pitch_rows + t) * 512;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c0 = 128 * q + 4 * lane;
      const float4 gg = *reinterpret_cast<const float4*>(&sg[c0]);
      const float4 bb = *reinterpret_cast<const float4*>(&sbe[c0]);
      const float ga[4] = {gg.x, gg.y, gg.z, gg.w}, ba[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int u2 = 0; u2 < 2; ++u2) {
        if (u2 == 1 && !two) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = (v[u2][4 * q + e] - mean[u2]) * rstd[u2] * ga[e] + ba[e];
        gelu_pair<T>(o[0], o[1]);
        gelu_pair<T>(o[2], o[3]);
        if (sizeof(T) == 2) {
          uint2 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
          h2[0] = __floats2bfloat162_rn(o[0], o[1]);
          h2[1] = __floats2bfloat162_rn(o[2], o[3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(yr) + u2 * 512 + c0) = u;
        } else {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(yr) + u2 * 512 + c0) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm rows (warp per row), optional pre-add and GELU
// Template R: 1 = the row unpacked to fp32 registers; 3 = the row kept in storage form (wide 16-bit rows, see the launcher);
// 2 = two rows per warp in storage form (measured no faster, not instantiated).
// Rows stay in registers in their STORAGE form (16 bytes per 8 values in the 16-bit modes) and are unpacked in each of the
// three passes (sum, centred sum of squares, normalise): 16 instead of 32 registers per 1024-channel row, so that two rows
// per warp still fit 64 registers and the SM keeps 32+ warps resident.
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void get(float (&v)[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
  __device__ __forceinline__ void add(const float* p) {
    const float4 c = *reinterpret_cast<const float4*>(p), d = *reinterpret_cast<const float4*>(p + 4);
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w; b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
  }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  // x + add is formed in fp32 and NOT rounded back in the one-row kernel; re-rounding here would change results, so the
  // pre-add case keeps using R = 1 with fp32 registers (see the launcher)
};

template <typename T, int CHUNKS, int R>
__global__ void __launch_bounds__(256, R == 2 ? 3 : (R == 3 ? 4 : 1))
layernorm_kernel(const T* __restrict__ x, long long rows, long long ldx, const T* __restrict__ add,
                 long long add_rows, const float* __restrict__ g, const float* __restrict__ b, float eps, int act,
                 T* __restrict__ y, long long ldy) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int C = CHUNKS * 256;
  const long long row0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (R == 2 ? 2 : 1);
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  if (R == 1) {
    float v[CHUNKS][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) Vec8<T>::load(x + row0 * ldx + j * 256 + lane * 8, v[j]);
    if (add) {
#pragma unroll
      for (int j = 0; j < CHUNKS; ++j) {
        float a[8];
        Vec8<T>::load(add + (row0 % add_rows) * C + j * 256 + lane * 8, a);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] += a[e];
      }
    }
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[j][e];
    const float mean = warp_sum(s) * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[j][e] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / (float)C) + eps);
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      const int c0 = j * 256 + lane * 8;
      float gg[8], bb[8], o[8];
      Vec8<float>::load(g + c0, gg);
      Vec8<float>::load(b + c0, bb);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * gg[e] + bb[e];
      if (act == ACT_GELU_L) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) gelu_pair<T>(o[e], o[e + 1]);
      }
      Vec8<T>::store(y + row0 * ldy + c0, o);
    }
    return;
  }
  // R == 2 / 3 (3 = ONE row, storage form, 4 blocks per SM), no pre-add: same arithmetic per row as above
  constexpr int NR = R == 2 ? 2 : 1;
  Raw8<T> raw[NR][CHUNKS];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;          // a ragged last pair re-reads the last row
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) raw[r][j].load(x + row * ldx + j * 256 + lane * 8);
  }
  float mean[NR], rstd[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      float v[8];
      raw[r][j].get(v);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[e];
    }
    mean[r] = warp_sum(s) * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < CHUNKS; ++j) {
      float v[8];
      raw[r][j].get(v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[e] - mean[r]; q = fmaf(d, d, q); }
    }
    rstd[r] = rsqrtf(warp_sum(q) * (1.0f / (float)C) + eps);
  }
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j) {
    const int c0 = j * 256 + lane * 8;
    float gg[8], bb[8];
    Vec8<float>::load(g + c0, gg);
    Vec8<float>::load(b + c0, bb);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (row0 + r < rows) {
        float v[8], o[8];
        raw[r][j].get(v);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean[r]) * rstd[r] * gg[e] + bb[e];
        if (act == ACT_GELU_L) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) gelu_pair<T>(o[e], o[e + 1]);
        }
        Vec8<T>::store(y + (row0 + r) * ldy + c0, o);
      }
    }
  }
}

template <typename T>
__global__ void add_rows_kernel(const T* __restrict__ x, long long rows, int c, const T* __restrict__ add,
                                long long add_rows, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = c / 8;
  if (i >= rows * c8) return;
  const long long r = i / c8;
  const int cc = (int)(i % c8) * 8;
  float a[8], bb[8];
  Vec8<T>::load(x + r * c + cc, a);
  Vec8<T>::load(add + (r % add_rows) * c + cc, bb);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] += bb[e];
  Vec8<T>::store(y + r * c + cc, a);
}

// ------------------------------------------------------------------ pixel subsampling for the stride-s 1x1 convolutions
// (architectures/video.py:13-15, 141-148: conv1 and the projection shortcut of the first block of layer2-4 both read the
// block input at every s-th pixel).  The sampled pixels are written once as rows of a wider [n*ho*wo, out_pitch] matrix
// (16 bytes per thread), next to which conv2 later places its output, so that conv1 becomes a plain row GEMM and
// conv3 + shortcut one K-concatenated GEMM.
__global__ void subsample_rows_kernel(const uint4* __restrict__ x, int n, int h, int w, int c16, int stride, int ho, int wo,
                                      uint4* __restrict__ y, long long y_pitch16) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * ho * wo * c16;
  if (i >= total) return;
  const int cc = (int)(i % c16);
  const long long m = i / c16;
  const int ow = (int)(m % wo);
  const int oh = (int)((m / wo) % ho);
  const long long img = m / ((long long)wo * ho);
  y[m * y_pitch16 + cc] = __ldg(x + ((img * h + (long long)oh * stride) * w + (long long)ow * stride) * c16 + cc);
}

// ------------------------------------------------------------------ multi-head self-attention, T tokens, no mask
// (HF Wav2Vec2Attention eval path; attention_layers.py:10-38).  One CTA per (window, head, 32-query
// tile); K and V of the head live in shared memory as fp32 (K rows padded by 1 float so that the
// per-lane key reads are bank-conflict free); one warp finishes one query at a time.
template <typename T, int DH>
__global__ void __launch_bounds__(256)
attention_kernel(const T* __restrict__ qkv, int t, int heads, float scale, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* sk = sm;                                  // [t][DH+1]
  float* sv = sm + (size_t)t * (DH + 1);           // [t][DH]
  float* sq = sv + (size_t)t * DH;                 // [8 warps][DH]
  float* sp = sq + 8 * DH;                         // [8 warps][t rounded to 32]
  const int tp = (t + 31) & ~31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row_stride = 3ll * heads * DH;
  const T* base = qkv + (long long)b * t * row_stride;
  for (int i = threadIdx.x; i < t * DH; i += blockDim.x) {
    const int j = i / DH, d = i % DH;
    sk[j * (DH + 1) + d] = to_f32(base[j * row_stride + (long long)(heads + head) * DH + d]);
    sv[j * DH + d] = to_f32(base[j * row_stride + (long long)(2 * heads + head) * DH + d]);
  }
  __syncthreads();
  const int q_end = min(t, (int)(blockIdx.x + 1) * 32);
  for (int qi = blockIdx.x * 32 + wid; qi < q_end; qi += 8) {
    for (int d = lane; d < DH; d += 32) sq[wid * DH + d] = to_f32(base[qi * row_stride + (long long)head * DH + d]) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < tp; j += 32) {
      float s = -INFINITY;
      if (j < t) {
        s = 0.f;
#pragma unroll 8
        for (int d = 0; d < DH; ++d) s = fmaf(sq[wid * DH + d], sk[j * (DH + 1) + d], s);
      }
      sp[wid * tp + j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < tp; j += 32) {
      const float e = j < t ? expf(sp[wid * tp + j] - mx) : 0.f;
      sp[wid * tp + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < DH; d += 32) {
      float o = 0.f;
      for (int j = 0; j < t; ++j) o = fmaf(sp[wid * tp + j], sv[j * DH + d], o);
      out[((long long)b * t + qi) * heads * DH + head * DH + d] = from_f32<T>(o * inv);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ audio head pools (audio_8_cl.py:146-159)
template <typename T>
__global__ void maxpool1d5_relu_kernel(const T* __restrict__ x, int n, int t, int c, int to, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * to * c) return;
  const int cc = (int)(i % c);
  const int ot = (int)((i / c) % to);
  const int b = (int)(i / ((long long)c * to));
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float v = to_f32(x[((long long)b * t + ot * 5 + k) * c + cc]);
    m = (v > m || v != v) ? v : m;
  }
  y[i] = from_f32<T>(m < 0.f ? 0.f : m);
}
template <typename T>
__global__ void avgpool1d_relu_kernel(const T* __restrict__ x, int n, int t, int c, T* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * c) return;
  const int cc = (int)(i % c);
  const int b = (int)(i / c);
  float s = 0.f;
  for (int k = 0; k < t; ++k) s += to_f32(x[((long long)b * t + k) * c + cc]);
  const float a = s / (float)t;
  y[i] = from_f32<T>(a < 0.f ? 0.f : a);
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ x, long long n, TD* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = from_f32<TD>(to_f32(x[i]));
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


extern "C" int avcer_maxpool3x3s2(const void* x, int n, int h, int w, int c, void* y, int dtype, void* stream) {
  AVCER_REQUIRE(c % 8 == 0 && h >= 3 && w >= 3, "maxpool3x3s2: bad shape");
  const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(maxpool3x3s2_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), 
                            (const T*)x, n, h, w, c, ho, wo, (T*)y)));
  return check_launch("maxpool3x3s2");
}

extern "C" int avcer_avgpool(const void* x, int n, int hw, int c, void* y, int dtype, void* stream) {
  AVCER_REQUIRE(c % 8 == 0 && hw > 0, "avgpool: bad shape");
  const long long total = (long long)n * (c / 8);
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(avgpool_kernel<T>, blocks_for(total, 128), 128, 0, as_stream(stream), (const T*)x, n, hw, c, (T*)y)));
  return check_launch("avgpool");
}

extern "C" int avcer_small_linear(const void* x, int64_t n, int k, int64_t ldx, const float* w, const float* b, int m,
                                  int softmax, float* y, int dtype, void* stream) {
  AVCER_REQUIRE(m >= 1 && m <= 8 && k > 0, "small_linear: m must be in [1,8]");
  if (n == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(small_linear_kernel<T>, blocks_for(n * 32, 256), 256, 0, as_stream(stream), 
                            (const T*)x, n, k, ldx, w, b, m, softmax, y)));
  return check_launch("small_linear");
}

extern "C" int avcer_lstm_cell(const float* xproj, const int32_t* xidx, const float* hproj, float* c, void* h_out,
                               int64_t ldh, int64_t n, int hidden, int first, int split, float* h_f32, int dtype, void* stream) {
  AVCER_REQUIRE(hidden > 0 && ldh >= (split ? 3 : 1) * (int64_t)hidden, "lstm_cell: bad shape");
  AVCER_REQUIRE(!split || dtype == AVCER_BF16, "lstm_cell: the [hi | lo | hi] split output is a bf16 feature");
  const long long total = n * hidden;
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(lstm_cell_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), 
                            xproj, xidx, hproj, c, (T*)h_out, ldh, n, hidden, first, split, h_f32)));
  return check_launch("lstm_cell");
}

extern "C" int avcer_gru_cell(const float* xg, int64_t x_row0, int64_t x_row_stride, const float* hg, float* h_state, void* h_out,
                              int64_t ldh, int split, void* y, int64_t y_row0, int64_t y_row_stride, int64_t n, int hidden, int dtype,
                              void* stream) {
  AVCER_REQUIRE(hidden > 0 && ldh >= (split ? 3 : 1) * (int64_t)hidden, "gru_cell: bad shape");
  AVCER_REQUIRE(!split || dtype == AVCER_BF16, "gru_cell: the [hi | lo | hi] split output is a bf16 feature");
  AVCER_REQUIRE(xg != nullptr && hg != nullptr && h_state != nullptr && h_out != nullptr, "gru_cell: null pointer");
  const long long total = n * hidden;
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(gru_cell_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), xg, (long long)x_row0,
                                    (long long)x_row_stride, hg, h_state, (T*)h_out, (long long)ldh, split, (T*)y, (long long)y_row0,
                                    (long long)y_row_stride, (long long)n, hidden)));
  return check_launch("gru_cell");
}

extern "C" int avcer_split_bf16x3(const float* x, int64_t rows, int k, int64_t ldx, void* out, int64_t ldo, void* stream) {
  AVCER_REQUIRE(k > 0 && ldx >= k && ldo >= 3 * (int64_t)k, "split_bf16x3: bad shape");
  const long long total = rows * k;
  if (total == 0) return 0;
  launch_pdl(split_bf16x3_kernel, blocks_for(total, 256), 256, 0, as_stream(stream), x, (long long)rows, k, (long long)ldx,
             (__nv_bfloat16*)out, (long long)ldo);
  return check_launch("split_bf16x3");
}

extern "C" int avcer_audio_normalize_windows(const float* wav, const int64_t* starts, const int64_t* ends, int n_win,
                                             int win, int pad_mode, float* out, void* stream) {
  AVCER_REQUIRE(pad_mode >= 0 && pad_mode <= 2 && win > 0, "audio_normalize_windows: bad arguments");
  if (n_win == 0) return 0;
  launch_pdl(audio_normalize_kernel, n_win, 512, 0, as_stream(stream), wav, (const long long*)starts, (const long long*)ends, win, pad_mode, out);
  return check_launch("audio_normalize_windows");
}

extern "C" int avcer_w2v_conv0_ln_gelu(const float* x, int n, int t_in, const float* w, const float* b,
                                       const float* ln_g, const float* ln_b, void* y, int64_t y_pitch_rows, int dtype,
                                       void* stream) {
  AVCER_REQUIRE(t_in >= 10, "w2v_conv0: t_in too small");
  const int t_out = (t_in - 10) / 5 + 1;
  AVCER_REQUIRE(y_pitch_rows >= t_out, "w2v_conv0: y pitch too small");
  if (n == 0) return 0;
  const long long pairs = (long long)n * ((t_out + 1) / 2);
  long long gx = (pairs + 7) / 8;
  const long long cap = 3ll * num_sms();          // persistent: three resident blocks per SM
  if (gx > cap) gx = cap;
  AVCER_DISPATCH(dtype, (launch_pdl(w2v_conv0_kernel<T>, (unsigned)gx, 256, 0, as_stream(stream), x, n, t_in, t_out, w, b, ln_g, ln_b,
                                                                                    (T*)y, y_pitch_rows)));
  return check_launch("w2v_conv0");
}

extern "C" int avcer_layernorm(const void* x, int64_t rows, int c, int64_t ldx, const void* add, int64_t add_rows,
                               const float* g, const float* b, float eps, int act, void* y, int64_t ldy, int dtype,
                               void* stream) {
  AVCER_REQUIRE(c % 256 == 0 && c <= 1024, "layernorm: c must be a multiple of 256, at most 1024");
  AVCER_REQUIRE(add == nullptr || add_rows > 0, "layernorm: add_rows must be > 0 with add");
  if (rows == 0) return 0;
  // variant 1: the row in fp32 registers (pre-add, fp32 mode, narrow rows); variant 3: wide 16-bit rows without pre-add held
  // in storage form at 64 registers per thread -> 32 instead of 16 resident warps per SM (the 1024-channel kernel needed 119
  // registers).  Same arithmetic in the same order: bit-identical outputs; measured 18.3 -> 15.9 us per [12736, 1024] pass
  // with the input in L2 (scripts/time_ln.py).  Two rows per warp (variant 2) measured the same and was dropped.
  const int variant = (add == nullptr && dtype == AVCER_BF16 && c >= 768) ? 3 : 1;
#define AVCER_LN_CASE(ch, var)                                                                                    \
  if (c == ch * 256 && variant == var) {                                                                          \
    AVCER_DISPATCH(dtype, (launch_pdl(layernorm_kernel<T, ch, var>, blocks_for(rows * 32, 256), 256, 0, as_stream(stream), \
                              (const T*)x, rows, ldx, (const T*)add, add_rows > 0 ? add_rows : 1, g, b, eps, act, \
                              (T*)y, ldy)));                                                                      \
  }
  AVCER_LN_CASE(1, 1) AVCER_LN_CASE(2, 1) AVCER_LN_CASE(3, 1) AVCER_LN_CASE(4, 1) AVCER_LN_CASE(3, 3) AVCER_LN_CASE(4, 3)
#undef AVCER_LN_CASE
  return check_launch("layernorm");
}

extern "C" int avcer_subsample_rows(const void* x, int n, int h, int w, int c, int stride, void* y, int64_t y_pitch, int dtype,
                                    void* stream) {
  const int esz = dtype == AVCER_BF16 ? 2 : 4;
  AVCER_REQUIRE(dtype == AVCER_BF16 || dtype == AVCER_F32, "subsample_rows: unknown dtype %d", dtype);
  AVCER_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && stride >= 1, "subsample_rows: bad shape");
  AVCER_REQUIRE((c * esz) % 16 == 0 && (y_pitch * esz) % 16 == 0 && y_pitch >= c, "subsample_rows: rows must be multiples of 16 bytes");
  AVCER_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "subsample_rows: 16-byte alignment");
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const int c16 = c * esz / 16;
  const long long total = (long long)n * ho * wo * c16;
  if (total == 0) return 0;
  launch_pdl(subsample_rows_kernel, blocks_for(total, 256), 256, 0, as_stream(stream), (const uint4*)x, n, h, w, c16, stride, ho, wo,
             (uint4*)y, (long long)(y_pitch * esz / 16));
  return check_launch("subsample_rows");
}

extern "C" int avcer_add_rows(const void* x, int64_t rows, int c, const void* add, int64_t add_rows, void* y,
                              int dtype, void* stream) {
  AVCER_REQUIRE(c % 8 == 0 && add_rows > 0, "add_rows: bad shape");
  const long long total = rows * (c / 8);
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(add_rows_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), 
                            (const T*)x, rows, c, (const T*)add, add_rows, (T*)y)));
  return check_launch("add_rows");
}

template <typename T, int DH>
static int launch_attention(const void* qkv, int n, int t, int heads, float scale, void* out, cudaStream_t st) {
  const int tp = (t + 31) & ~31;
  const size_t smem = ((size_t)t * (DH + 1) + (size_t)t * DH + 8 * DH + 8 * tp) * sizeof(float);
  AVCER_REQUIRE(smem <= 220 * 1024, "attention: T=%d too long for the shared-memory kernel", t);
  auto kern = attention_kernel<T, DH>;
  AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((t + 31) / 32, heads, n);
  kern<<<grid, 256, smem, st>>>((const T*)qkv, t, heads, scale, (T*)out);
  return check_launch("attention");
}

namespace avcer {
int attention_tc(const void* qkv, int n, int t, int heads, int dh, float scale, void* out, cudaStream_t st);   // attention_tc.cu
}

extern "C" int avcer_attention(const void* qkv, int n, int t, int heads, int dh, float scale, void* out, int dtype,
                               void* stream) {
  AVCER_REQUIRE(dh == 32 || dh == 64, "attention: head dim must be 32 or 64");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == AVCER_BF16) return attention_tc(qkv, n, t, heads, dh, scale, out, st);
  if (dtype == AVCER_F32) return dh == 64 ? launch_attention<float, 64>(qkv, n, t, heads, scale, out, st)
                                          : launch_attention<float, 32>(qkv, n, t, heads, scale, out, st);
  return set_error("attention: unknown dtype %d", dtype);
}

extern "C" int avcer_maxpool1d5_relu(const void* x, int n, int t, int c, void* y, int dtype, void* stream) {
  const int to = t / 5;
  const long long total = (long long)n * to * c;
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(maxpool1d5_relu_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), 
                            (const T*)x, n, t, c, to, (T*)y)));
  return check_launch("maxpool1d5_relu");
}

extern "C" int avcer_avgpool1d_relu(const void* x, int n, int t, int c, void* y, int dtype, void* stream) {
  const long long total = (long long)n * c;
  if (total == 0) return 0;
  AVCER_DISPATCH(dtype, (launch_pdl(avgpool1d_relu_kernel<T>, blocks_for(total, 256), 256, 0, as_stream(stream), 
                            (const T*)x, n, t, c, (T*)y)));
  return check_launch("avgpool1d_relu");
}

extern "C" int avcer_cast(const void* x, int64_t n, int src_dtype, void* y, int dst_dtype, void* stream) {
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const unsigned g = blocks_for(n, 256);
  if (src_dtype == AVCER_F32 && dst_dtype == AVCER_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)x, n, (bf16*)y);
  else if (src_dtype == AVCER_BF16 && dst_dtype == AVCER_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)x, n, (float*)y);
  else if (src_dtype == AVCER_F32 && dst_dtype == AVCER_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)x, n, (float*)y);
  else if (src_dtype == AVCER_BF16 && dst_dtype == AVCER_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)x, n, (bf16*)y);
  else return set_error("cast: unknown dtypes %d -> %d", src_dtype, dst_dtype);
  return check_launch("cast");
}
