// K1: face-crop preprocessing (reference: src/data/utils.py:19-39 pth_processing, called per frame
// at src/get_prob_video.py:95-99, followed by the H2D copy at :108).
//
//   PIL resize((224,224), NEAREST) -> CHW u8 -> float32 -> channel flip (undoes the BGR2RGB of
//   get_prob_video.py:97, so the tensor is in cv2's BGR byte order) -> subtract
//   [91.4953, 103.8827, 131.0912] -> (no /255).
//
// The nearest-neighbour source index follows Pillow's affine-scale loop exactly: with
// a = in/224.0 (double), o = a*0.5, idx[x] = (int)o, o += a  -- an incremental double
// accumulation (not floor((x+.5)*a)), restated in resize_maps_kernel.
//
// Memory-bound: one CTA converts 4 output rows of one crop; source rows are staged through shared
// memory with 16-byte loads and every global store is a fully coalesced 16-byte vector.
#include "common.h"
#include <stdlib.h>

namespace avcer {

constexpr int OUT = 224;
constexpr int PADW = 240;     // padded row pitch (pixels) of layouts 1/2: 1920 B = 15 x 128 B per bf16 row
constexpr int PADH = 232;
constexpr int PAD0 = 2;       // TF-"same" leading pad of the 7x7/2 stem (video.py:65-81)
constexpr int ROWS = 4;       // output rows per CTA
constexpr int MAX_STAGE_W = 1024;

__constant__ float c_mean[3] = {91.4953f, 103.8827f, 131.0912f};

__global__ void resize_maps_kernel(const int* __restrict__ src_h, const int* __restrict__ src_w, int n,
                                   short* __restrict__ maps /*[n][2][224]: y map then x map*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 2) return;
  const int crop = i >> 1, axis = i & 1;
  const int in = axis == 0 ? src_h[crop] : src_w[crop];
  const double a = (double)in / 224.0;
  double o = __dmul_rn(a, 0.5);
  short* m = maps + (size_t)crop * 2 * OUT + axis * OUT;
  for (int x = 0; x < OUT; ++x) {
    int v = (int)o;
    if (v >= in) v = in - 1;          // Pillow skips such pixels; cannot happen for in >= 1
    m[x] = (short)v;
    o = __dadd_rn(o, a);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ offs, const int* __restrict__ src_h,
                  const int* __restrict__ src_w, const short* __restrict__ maps, void* __restrict__ dst) {
  __shared__ __align__(16) uint8_t rows[ROWS][MAX_STAGE_W * 3];
  __shared__ short xmap[OUT];
  const int crop = blockIdx.y;
  const int y0 = blockIdx.x * ROWS;
  const int sw = src_w ? src_w[crop] : OUT;
  const int sh = src_h ? src_h[crop] : OUT;
  (void)sh;
  const uint8_t* base = src + (offs ? offs[crop] : (long long)crop * OUT * OUT * 3);
  const short* cm = maps ? maps + (size_t)crop * 2 * OUT : nullptr;
  const bool staged = sw <= MAX_STAGE_W;
  const int row_bytes = sw * 3;

  if (cm) for (int x = threadIdx.x; x < OUT; x += blockDim.x) xmap[x] = cm[OUT + x];
  if (staged) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int sy = cm ? cm[y0 + r] : (y0 + r);
      const uint8_t* rp = base + (long long)sy * row_bytes;
      if (((reinterpret_cast<uintptr_t>(rp) & 15) == 0) && (row_bytes % 16 == 0)) {
        const uint4* rp4 = reinterpret_cast<const uint4*>(rp);
        uint4* sp4 = reinterpret_cast<uint4*>(rows[r]);
        for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) sp4[i] = __ldg(rp4 + i);
      } else {
        for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) rows[r][i] = __ldg(rp + i);
      }
    }
  }
  __syncthreads();

  auto fetch = [&](int r, int x, float (&v)[3]) {
    const int sx = cm ? xmap[x] : x;
    if (staged) {
      const uint8_t* px = &rows[r][sx * 3];
      v[0] = (float)px[0] - c_mean[0]; v[1] = (float)px[1] - c_mean[1]; v[2] = (float)px[2] - c_mean[2];
    } else {
      const int sy = cm ? cm[y0 + r] : (y0 + r);
      const uint8_t* px = base + ((long long)sy * sw + sx) * 3;
      v[0] = (float)__ldg(px) - c_mean[0]; v[1] = (float)__ldg(px + 1) - c_mean[1]; v[2] = (float)__ldg(px + 2) - c_mean[2];
    }
  };

  if (LAYOUT == 1) {
    // bf16 NHWC4, zero border: two pixels per 16-byte store
    __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst) + (size_t)crop * PADH * PADW * 4;
    for (int t = threadIdx.x; t < ROWS * (OUT / 2); t += blockDim.x) {
      const int r = t / (OUT / 2), q = t % (OUT / 2);
      float a[3], b[3];
      fetch(r, 2 * q, a);
      fetch(r, 2 * q + 1, b);
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
      h2[0] = __floats2bfloat162_rn(a[0], a[1]);
      h2[1] = __floats2bfloat162_rn(a[2], 0.f);
      h2[2] = __floats2bfloat162_rn(b[0], b[1]);
      h2[3] = __floats2bfloat162_rn(b[2], 0.f);
      uint4* o = reinterpret_cast<uint4*>(d + ((size_t)(y0 + r + PAD0) * PADW + PAD0 + 2 * q) * 4);
      __stcs(o, u);
      // full 32-byte sectors at both ends of the row (see preprocess_identity_kernel): the neighbouring border pixels are zero
      if (q == 0 || q == OUT / 2 - 1) __stcs(q == 0 ? o - 1 : o + 1, make_uint4(0u, 0u, 0u, 0u));
    }
  } else if (LAYOUT == 2) {
    float* d = static_cast<float*>(dst) + (size_t)crop * PADH * PADW * 4;
    for (int t = threadIdx.x; t < ROWS * OUT; t += blockDim.x) {
      const int r = t / OUT, x = t % OUT;
      float a[3];
      fetch(r, x, a);
      *reinterpret_cast<float4*>(d + ((size_t)(y0 + r + PAD0) * PADW + PAD0 + x) * 4) = make_float4(a[0], a[1], a[2], 0.f);
    }
  } else {
    // fp32 NCHW: the reference tensor itself
    float* d = static_cast<float*>(dst) + (size_t)crop * 3 * OUT * OUT;
    for (int t = threadIdx.x; t < ROWS * OUT * 3; t += blockDim.x) {
      const int c = t / (ROWS * OUT), rem = t % (ROWS * OUT);
      const int r = rem / OUT, x = rem % OUT;
      float a[3];
      fetch(r, x, a);
      d[(size_t)c * OUT * OUT + (size_t)(y0 + r) * OUT + x] = a[c];
    }
  }
}

// Fast path for packed 224x224 crops (resize = identity; BASELINE configs): one CTA converts 16 rows.
// 10752 contiguous input bytes are staged with three 16-byte loads in flight per thread, then every
// thread emits seven fully coalesced 16-byte stores (two bf16 NHWC4 pixels each).
template <int LAYOUT, int ROWS_I = 16, int CS = 0, int FULL_SECTORS = 1>
__global__ void __launch_bounds__(256)
preprocess_identity_kernel(const uint8_t* __restrict__ src, void* __restrict__ dst) {
  __shared__ __align__(16) uint8_t sm[ROWS_I * OUT * 3];
  const int crop = blockIdx.y;
  const int y0 = blockIdx.x * ROWS_I;
  const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)crop * OUT * OUT * 3 + (size_t)y0 * OUT * 3);
  constexpr int NV = ROWS_I * OUT * 3 / 16;      // 672 for 16 rows
  constexpr int NL = (NV + 255) / 256;
  uint4 tmp[NL];
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < NV) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(tmp[k].x), "=r"(tmp[k].y), "=r"(tmp[k].z), "=r"(tmp[k].w) : "l"(g + i));
  }
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < NV) reinterpret_cast<uint4*>(sm)[i] = tmp[k];
  }
  __syncthreads();
  const unsigned short* s16 = reinterpret_cast<const unsigned short*>(sm);
  if (LAYOUT == 1) {
    __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst) + (size_t)crop * PADH * PADW * 4;
#pragma unroll
    for (int it = 0; it < (ROWS_I * (OUT / 2) + 255) / 256; ++it) {
      const int t = threadIdx.x + it * 256;
      if (t >= ROWS_I * (OUT / 2)) break;
      const int r = t / (OUT / 2), q = t % (OUT / 2);
      const unsigned short* px = s16 + r * (OUT * 3 / 2) + q * 3;        // 6 bytes = 2 BGR pixels
      const unsigned a = px[0], b = px[1], c = px[2];
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
      h2[0] = __floats2bfloat162_rn((float)(a & 0xff) - c_mean[0], (float)(a >> 8) - c_mean[1]);
      h2[1] = __floats2bfloat162_rn((float)(b & 0xff) - c_mean[2], 0.f);
      h2[2] = __floats2bfloat162_rn((float)(b >> 8) - c_mean[0], (float)(c & 0xff) - c_mean[1]);
      h2[3] = __floats2bfloat162_rn((float)(c >> 8) - c_mean[2], 0.f);
      uint4* o = reinterpret_cast<uint4*>(d + ((size_t)(y0 + r + PAD0) * PADW + PAD0 + 2 * q) * 4);
      if (CS) __stcs(o, u); else *o = u;
      if (FULL_SECTORS && (q == 0 || q == OUT / 2 - 1)) {
        // The interior of a row starts 16 bytes into a 32-byte sector and ends 16 bytes into another one (2 border pixels +
        // 224 pixels of 8 bytes): two PARTIAL sector writes per row, each a read-modify-write at the DRAM.  Rewriting the two
        // neighbouring (zero) border pixels makes every sector of the row a full write: 116 -> 87 us per 1024 crops.
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (CS) __stcs(q == 0 ? o - 1 : o + 1, z); else *(q == 0 ? o - 1 : o + 1) = z;
      }
    }
  } else {
    float* d = static_cast<float*>(dst) + (size_t)crop * PADH * PADW * 4;
    for (int t = threadIdx.x; t < ROWS_I * OUT; t += 256) {
      const int r = t / OUT, x = t % OUT;
      const uint8_t* px = sm + (r * OUT + x) * 3;
      *reinterpret_cast<float4*>(d + ((size_t)(y0 + r + PAD0) * PADW + PAD0 + x) * 4) =
          make_float4((float)px[0] - c_mean[0], (float)px[1] - c_mean[1], (float)px[2] - c_mean[2], 0.f);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Audio decode seam (reference: src/data/utils.py:49-60 convert_mp4_to_mp3 after its ffmpeg
// call): interleaved int16 PCM -> float (1/32768) -> channel mean -> torchaudio's polyphase sinc resampler
// (transforms.Resample defaults: a strided conv1d of the zero-padded signal with `nnew` filters of `taps` =
// 2*width + orig coefficients; output sample f*nnew + j = sum_k bank[j][k] * x[f*orig + k - width]).
// One CTA produces RS_FRAMES frames (RS_FRAMES*nnew outputs): the input span is converted once into shared memory,
// thread j walks the taps with the bank stored tap-major ([taps][nnew]) so that a warp reads consecutive coefficients.
constexpr int RS_FRAMES = 4;

__device__ __forceinline__ float pcm_mono(const short* __restrict__ pcm, long long i, long long n, int ch) {
  if (i < 0 || i >= n) return 0.f;
  float s = 0.f;
  for (int c = 0; c < ch; ++c) s += (float)pcm[i * ch + c] * (1.0f / 32768.0f);
  return ch > 1 ? s / (float)ch : s;
}

__global__ void __launch_bounds__(256)
pcm16_resample_kernel(const short* __restrict__ pcm, long long n, int ch, const float* __restrict__ bank, int orig, int nnew,
                      int width, long long n_out, float* __restrict__ out) {
  extern __shared__ float xs[];
  const int taps = 2 * width + orig;
  const long long f0 = (long long)blockIdx.x * RS_FRAMES;
  const int span = (RS_FRAMES - 1) * orig + taps;
  const long long base = f0 * orig - width;
  for (int i = threadIdx.x; i < span; i += blockDim.x) xs[i] = pcm_mono(pcm, base + i, n, ch);
  __syncthreads();
  for (int j = threadIdx.x; j < nnew; j += blockDim.x) {
    float acc[RS_FRAMES];
#pragma unroll
    for (int f = 0; f < RS_FRAMES; ++f) acc[f] = 0.f;
    for (int k = 0; k < taps; ++k) {
      const float c = __ldg(bank + (size_t)k * nnew + j);
#pragma unroll
      for (int f = 0; f < RS_FRAMES; ++f) acc[f] = fmaf(c, xs[f * orig + k], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < RS_FRAMES; ++f) {
      const long long o = (f0 + f) * nnew + j;
      if (o < n_out) out[o] = acc[f];
    }
  }
}

__global__ void pcm16_mono_kernel(const short* __restrict__ pcm, long long n, int ch, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = pcm_mono(pcm, i, n, ch);
}

}  // namespace avcer

using namespace avcer;

extern "C" int avcer_preprocess_maps(const int32_t* src_h, const int32_t* src_w, int n, int16_t* maps, void* stream) {
  AVCER_REQUIRE(n >= 0, "preprocess_maps: negative n");
  if (n == 0) return 0;
  resize_maps_kernel<<<(2 * n + 63) / 64, 64, 0, as_stream(stream)>>>(src_h, src_w, n, maps);
  return check_launch("resize_maps_kernel");
}

extern "C" int avcer_preprocess_u8(const uint8_t* src, const int64_t* src_offsets, const int32_t* src_h,
                                   const int32_t* src_w, const int16_t* maps, int n, void* dst, int dst_layout,
                                   void* stream) {
  AVCER_REQUIRE(n >= 0, "preprocess: negative n");
  AVCER_REQUIRE(dst_layout >= 0 && dst_layout <= 2, "preprocess: unknown layout %d", dst_layout);
  AVCER_REQUIRE((src_h == nullptr) == (src_w == nullptr) && (src_h == nullptr) == (maps == nullptr),
                "preprocess: src_h, src_w and maps must be given together (or all NULL for packed 224x224 crops)");
  if (n == 0) return 0;
  dim3 grid(OUT / ROWS, n);
  AVCER_REQUIRE(n <= 65535, "preprocess: at most 65535 crops per call");
  cudaStream_t st = as_stream(stream);
  const long long* offs = reinterpret_cast<const long long*>(src_offsets);
  if (src_offsets == nullptr && src_h == nullptr && dst_layout != 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // AVCER_K1_STREAM=0: plain stores instead of streaming (st.global.cs) ones.  At 1024 crops: 116 us (0.62 of the copy peak
    // algorithmic) originally; 109 us with streaming stores; 87 us (0.82) once the two partial 32-byte sectors per row were
    // turned into full writes (see the kernel); 8- or 32-row tiles instead of 16 made no difference.
    static const int stream_stores = getenv("AVCER_K1_STREAM") ? atoi(getenv("AVCER_K1_STREAM")) : 1;
    if (dst_layout == 1) {
      static const int partial = getenv("AVCER_K1_PARTIAL") ? atoi(getenv("AVCER_K1_PARTIAL")) : 0;   // 1: the round-1 kernel (evidence runs)
      if (partial) preprocess_identity_kernel<1, 16, 0, 0><<<dim3(OUT / 16, n), 256, 0, st>>>(src, dst);
      else if (stream_stores) preprocess_identity_kernel<1, 16, 1><<<dim3(OUT / 16, n), 256, 0, st>>>(src, dst);
      else preprocess_identity_kernel<1, 16, 0><<<dim3(OUT / 16, n), 256, 0, st>>>(src, dst);
    } else {
      preprocess_identity_kernel<2><<<dim3(OUT / 16, n), 256, 0, st>>>(src, dst);
    }
    return check_launch("preprocess_identity_kernel");
  }
  if (dst_layout == 0) preprocess_kernel<0><<<grid, 256, 0, st>>>(src, offs, src_h, src_w, maps, dst);
  else if (dst_layout == 1) preprocess_kernel<1><<<grid, 256, 0, st>>>(src, offs, src_h, src_w, maps, dst);
  else preprocess_kernel<2><<<grid, 256, 0, st>>>(src, offs, src_h, src_w, maps, dst);
  return check_launch("preprocess_kernel");
}

extern "C" int avcer_pcm16_resample(const int16_t* pcm, int64_t n, int channels, const float* bank_tap_major, int orig,
                                    int nnew, int width, float* out, int64_t n_out, void* stream) {
  AVCER_REQUIRE(n >= 0 && channels >= 1 && n_out >= 0, "pcm16_resample: bad sizes");
  if (n_out == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (bank_tap_major == nullptr) {                 // same rate: scale + channel mean only
    AVCER_REQUIRE(n_out == n, "pcm16_resample: n_out must equal n without a filter bank");
    pcm16_mono_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pcm, n, channels, out);
    return check_launch("pcm16_mono_kernel");
  }
  AVCER_REQUIRE(orig >= 1 && nnew >= 1 && width >= 0, "pcm16_resample: bad resampling ratio");
  const int64_t frames = n / orig + 1;             // conv1d output length over the padded signal
  AVCER_REQUIRE(n_out <= frames * nnew, "pcm16_resample: n_out %lld exceeds the %lld samples the filter produces", (long long)n_out,
                (long long)(frames * nnew));
  const int taps = 2 * width + orig;
  const size_t smem = ((size_t)(RS_FRAMES - 1) * orig + taps) * sizeof(float);
  AVCER_REQUIRE(smem <= 200 * 1024, "pcm16_resample: ratio %d/%d needs %zu B of shared memory", orig, nnew, smem);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    AVCER_CUDA(cudaFuncSetAttribute(pcm16_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int64_t used_frames = (n_out + nnew - 1) / nnew;
  const unsigned grid = (unsigned)((used_frames + RS_FRAMES - 1) / RS_FRAMES);
  pcm16_resample_kernel<<<grid, 256, smem, st>>>(pcm, (long long)n, channels, bank_tap_major, orig, nnew, width, (long long)n_out, out);
  return check_launch("pcm16_resample_kernel");
}
