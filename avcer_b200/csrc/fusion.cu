// K4 and the alignment glue around it: probability fusion -> compound-expression rule -> argmax
// (reference: src/run.py:105-165, src/data/utils.py:125-127,222-241), per-frame mean of audio
// window logits (run.py:90 groupby.mean + get_prob_audio_8_cl.py:94-101) and row gathers
// (carry-forward of get_prob_video.py:157-178, column permutation run.py:85-88).
//
// Bit-exactness contract of K4: the reference computes in numpy with IEEE binary64 (binary32
// when no weight matrix is given) using separate multiply and add instructions.  The kernel
// uses the *_rn intrinsics so the compiler can never contract them into FMAs.
#include "common.h"

namespace avcer {

// Compound expressions as pairs of basic-emotion indices in audio order (run.py:66-74).
// constexpr (not __constant__): the indices must fold at compile time so that the 7-value rows stay in registers
__host__ __device__ constexpr int pair_a(int k) { return k == 0 ? 3 : k == 1 ? 4 : k == 2 ? 5 : k == 3 ? 2 : k == 4 ? 1 : k == 5 ? 3 : 1; }
__host__ __device__ constexpr int pair_b(int k) { return k < 5 ? 6 : 5; }

struct FuseParams {
  double w[3][7];      // weights_1[m][c] (unused when !has_w1)
  double w2[3];        // weights_2[m]
  double ce_w[7][2];   // rule-2 pair weights (1,1 when !ce_weights_type)
  int has_w1, ce_mask;
  int w2_one, cew_one;   // every weights_2[m] == 1 / every rule weight == 1: the exact no-op multiplies are skipped
                         // (B200's vector FP64 rate, not HBM, bounds this kernel)
};

template <typename TC> struct Arith;
template <> struct Arith<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double div3(double a) { return __ddiv_rn(a, 3.0); }
  static __device__ __forceinline__ double thr() { return 1.0 / 7.0; }
};
template <> struct Arith<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float div3(float a) { return __fdiv_rn(a, 3.0f); }
  static __device__ __forceinline__ float thr() { return (float)(1.0 / 7.0); }
};

// numpy argmax over the 7 compound scores: first maximum, NaN is "largest".
template <typename TC>
__device__ __forceinline__ long long compound_argmax(const TC (&s)[7], const FuseParams& p) {
  using A = Arith<TC>;
  TC v[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) v[c] = p.ce_mask ? (s[c] > A::thr() ? s[c] : (TC)0) : s[c];
  TC best = (TC)0;
  int bi = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    TC a = v[pair_a(k)], b = v[pair_b(k)];
    if (!p.cew_one) {                      // x * 1 == x bit-for-bit (NaN, inf and -0 included)
      a = A::mul(a, (TC)p.ce_w[k][0]);
      b = A::mul(b, (TC)p.ce_w[k][1]);
    }
    const TC pr = A::add(a, b);
    if (k == 0) { best = pr; bi = 0; }
    // "pr > best, or pr is NaN and best is not" in two compares instead of three: !(pr <= best) is the unordered
    // greater-than (true when either side is NaN), and a NaN best never gives way
    else if (!(pr <= best) && best == best) { best = pr; bi = k; }
  }
  return bi;
}

constexpr int FUSE_THREADS = 128;

template <typename TIn, typename TC>
__device__ __forceinline__ void fuse_one(const TIn (&a)[7], const TIn (&b)[7], const TIn (&c)[7], const FuseParams& p,
                                         long long (&lab)[4]) {
  using A = Arith<TC>;
  TC x[3][7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { x[0][k] = (TC)a[k]; x[1][k] = (TC)b[k]; x[2][k] = (TC)c[k]; }
  TC av[7];
  if (p.has_w1) {
    // predictions[m] * weights_1[m] * weights_2[m], summed left to right (run.py:108-111)
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        x[m][k] = A::mul(x[m][k], (TC)p.w[m][k]);
        if (!p.w2_one) x[m][k] = A::mul(x[m][k], (TC)p.w2[m]);
      }
#pragma unroll
    for (int k = 0; k < 7; ++k) av[k] = A::add(A::add(x[0][k], x[1][k]), x[2][k]);
  } else {
    // np.sum(predictions, axis=0) / 3 (run.py:113-114)
#pragma unroll
    for (int k = 0; k < 7; ++k) av[k] = A::div3(A::add(A::add(x[0][k], x[1][k]), x[2][k]));
  }
  lab[0] = compound_argmax<TC>(av, p);
  lab[1] = compound_argmax<TC>(x[0], p);
  lab[2] = compound_argmax<TC>(x[1], p);
  lab[3] = compound_argmax<TC>(x[2], p);
}

// Warp-cooperative streaming pass.  A warp owns 32*FPT consecutive frames per iteration (FPT = 4 for f32, 2 for f64
// inputs): the three probability streams are copied global -> shared with fully coalesced 16-byte cp.async (one
// contiguous 3584-byte span per stream, no register staging), then lane l finishes frames l, l+32, ... of the chunk,
// reading its 7 values per stream with a 7-word lane stride (conflict-free) and writing its labels as coalesced 8-byte
// stores.  Keeping the chunk in shared memory instead of registers (the first version held 84 staging + 84 fp64
// registers per lane: 198 registers, 8 warps per SM, fp64 pipe 30 % busy, latency-bound) lets 5 blocks share an SM.
template <typename TIn, typename TC>
__global__ void __launch_bounds__(FUSE_THREADS, 5)
fuse_compound_kernel(const TIn* __restrict__ pvs, const TIn* __restrict__ pvd, const TIn* __restrict__ pa,
                     long long n, const FuseParams p, long long* __restrict__ labels, long long ld) {
  constexpr int FPT = 16 / sizeof(TIn);            // frames per lane and iteration
  constexpr int NV = 7;                            // 16-byte vectors per lane per stream
  constexpr int WARPS = FUSE_THREADS / 32;
  __shared__ uint4 tile[WARPS][3][32 * NV];        // 10.5 KB per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long groups = n / (32 * FPT);         // full warp-chunks
  const TIn* src[3] = {pvs, pvd, pa};
  for (long long g = (long long)blockIdx.x * WARPS + warp; g < groups; g += (long long)gridDim.x * WARPS) {
    const long long f0 = g * 32 * FPT;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const uint4* gsrc = reinterpret_cast<const uint4*>(src[m] + f0 * 7);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&tile[warp][m][lane + 32 * k]));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc + lane + 32 * k) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < FPT; ++j) {
      const int fi = j * 32 + lane;                // frame inside the chunk
      TIn a[7], b[7], c[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        a[k] = reinterpret_cast<const TIn*>(&tile[warp][0][0])[fi * 7 + k];
        b[k] = reinterpret_cast<const TIn*>(&tile[warp][1][0])[fi * 7 + k];
        c[k] = reinterpret_cast<const TIn*>(&tile[warp][2][0])[fi * 7 + k];
      }
      long long lab[4];
      fuse_one<TIn, TC>(a, b, c, p, lab);
#pragma unroll
      for (int s = 0; s < 4; ++s) labels[s * ld + f0 + fi] = lab[s];
    }
    __syncwarp();                                  // the chunk is consumed: the next cp.async may overwrite it
  }
  // tail frames (n % (32*FPT)) one per thread, by block 0
  const long long tail0 = groups * 32 * FPT;
  for (long long f = tail0 + threadIdx.x + (long long)blockIdx.x * blockDim.x; f < n; f += (long long)gridDim.x * blockDim.x) {
    TIn a[7], b[7], c[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = pvs[f * 7 + k]; b[k] = pvd[f * 7 + k]; c[k] = pa[f * 7 + k]; }
    long long lab[4];
    fuse_one<TIn, TC>(a, b, c, p, lab);
#pragma unroll
    for (int s = 0; s < 4; ++s) labels[s * ld + f] = lab[s];
  }
}

// Same arithmetic for buffers that are not 16-byte aligned (views into larger tensors): one frame per thread.
template <typename TIn, typename TC>
__global__ void __launch_bounds__(FUSE_THREADS)
fuse_compound_scalar_kernel(const TIn* __restrict__ pvs, const TIn* __restrict__ pvd, const TIn* __restrict__ pa,
                            long long n, const FuseParams p, long long* __restrict__ labels, long long ld) {
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += (long long)gridDim.x * blockDim.x) {
    TIn a[7], b[7], c[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = pvs[f * 7 + k]; b[k] = pvd[f * 7 + k]; c[k] = pa[f * 7 + k]; }
    long long lab[4];
    fuse_one<TIn, TC>(a, b, c, p, lab);
#pragma unroll
    for (int s = 0; s < 4; ++s) labels[s * ld + f] = lab[s];
  }
}

template <typename T>
__global__ void softmax7_kernel(const T* __restrict__ x, long long n, int ld, T* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T v[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) v[c] = x[i * ld + c];
  // np.max propagates NaN
  T m = v[0];
#pragma unroll
  for (int c = 1; c < 7; ++c) m = (v[c] > m || v[c] != v[c]) ? v[c] : m;
  T s = (T)0;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    v[c] = sizeof(T) == 4 ? (T)expf((float)(v[c] - m)) : (T)exp((double)(v[c] - m));
  }
  // np.sum over 7 contiguous elements: plain left-to-right (below the pairwise threshold)
#pragma unroll
  for (int c = 0; c < 7; ++c) s = s + v[c];
#pragma unroll
  for (int c = 0; c < 7; ++c) y[i * 7 + c] = v[c] / s;
}

// pandas group_mean: Kahan-compensated sum in the column dtype (float32 for the in-memory drivers' tables, float64 for
// tables read back from CSV, get_pred_av.py:246-249) in row order, NaN skipped, divided by the count in that dtype;
// nobs == 0 -> NaN.
template <typename T> struct KahanOps;
template <> struct KahanOps<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct KahanOps<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000ll); }
};
template <typename T>
__global__ void window_to_frame_mean_kernel(const T* __restrict__ logits, int n_win, int ncls,
                                            const int* __restrict__ f_lo, const int* __restrict__ f_hi,
                                            long long n_frames, T* __restrict__ out) {
  using K = KahanOps<T>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames * ncls) return;
  const long long f = i / ncls;
  const int c = (int)(i % ncls);
  // first window with f_hi > f (f_hi is non-decreasing)
  int lo = 0, hi = n_win;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((long long)f_hi[mid] > f) hi = mid; else lo = mid + 1;
  }
  T sum = 0, comp = 0, nobs = 0;
  for (int w = lo; w < n_win && (long long)f_lo[w] <= f; ++w) {
    if ((long long)f_hi[w] <= f) continue;
    const T val = logits[(long long)w * ncls + c];
    if (val == val) {
      nobs = K::add(nobs, (T)1);
      const T y = K::sub(val, comp);
      const T t = K::add(sum, y);
      comp = K::sub(K::sub(t, sum), y);
      if (comp != comp) comp = 0;
      sum = t;
    }
  }
  out[i] = nobs == (T)0 ? K::nan() : K::div(sum, nobs);
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, const int* __restrict__ idx, long long n_out,
                                   int ncols, const int* __restrict__ perm, T* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * ncols) return;
  const long long r = i / ncols;
  const int c = (int)(i % ncols);
  const int s = idx ? idx[r] : (int)r;
  out[i] = s >= 0 ? src[(long long)s * ncols + (perm ? perm[c] : c)] : (T)0;
}

// data/utils.py:222-241 as a stand-alone op (the reference exposes it as a function): arbitrary
// pair table, scores returned as float64 like the reference's np.zeros((n, k)) result array.
struct PairTable {
  int i1[16], i2[16];
  double w1[16], w2[16];
  int k;
};
template <typename TC>
__global__ void compound_scores_kernel(const TC* __restrict__ pred, long long n, int ncols, const PairTable t,
                                       int ce_mask, double* __restrict__ out) {
  using A = Arith<TC>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * t.k) return;
  const long long r = i / t.k;
  const int k = (int)(i % t.k);
  TC a = pred[r * ncols + t.i1[k]], b = pred[r * ncols + t.i2[k]];
  if (ce_mask) {
    a = a > A::thr() ? a : (TC)0;
    b = b > A::thr() ? b : (TC)0;
  }
  out[i] = (double)A::add(A::mul(a, (TC)t.w1[k]), A::mul(b, (TC)t.w2[k]));
}

// ------------------------------------------------------------------------------------------------
// Weight search: data/utils.py:138-163 (get_weights_prob_model: 10 000 Dirichlet
// weight sets) and :166-209 (get_weights_v_model / get_weights_av_model: grid over per-model weights).
// For every candidate weight set the reference recomputes  argmax_c sum_m P_m[f,c] * w[m,c]  over all frames
// and a classification report.  Here one launch evaluates all candidates: lanes own weight sets (their 21
// weights live in registers, their 7x7 confusion counters in a private shared-memory row -- no atomics), a
// tile of frames is staged in shared memory and broadcast to the lanes.  Arithmetic is the reference's:
// binary64, (p0*w0 + p1*w1) + p2*w2 with separate multiplies and adds, first-maximum argmax.
constexpr int WS_FRAMES = 128;      // frames per staged tile
constexpr int WS_WARPS = 8;         // 256 weight sets per CTA

template <int M>
__global__ void __launch_bounds__(WS_WARPS * 32)
weight_search_kernel(const double* __restrict__ preds, long long n, const int* __restrict__ gt,
                     const double* __restrict__ weights, long long n_weights, int frame_groups,
                     unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned char ws_smem[];
  double* sp = reinterpret_cast<double*>(ws_smem);                               // [WS_FRAMES][M*7]
  int* sgt = reinterpret_cast<int*>(sp + WS_FRAMES * M * 7);                    // [WS_FRAMES]
  unsigned* cnt = reinterpret_cast<unsigned*>(sgt + WS_FRAMES);                  // [256][49]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ws = (long long)blockIdx.x * (WS_WARPS * 32) + threadIdx.x;    // this lane's weight set
  const bool active = ws < n_weights;
  double w[M][7];
#pragma unroll
  for (int m = 0; m < M; ++m)
#pragma unroll
    for (int c = 0; c < 7; ++c) w[m][c] = active ? weights[(ws * M + m) * 7 + c] : 0.0;
  unsigned* my = cnt + threadIdx.x * 49;
  for (int i = 0; i < 49; ++i) my[i] = 0u;
  // frames of this CTA: contiguous slice blockIdx.y of `frame_groups`
  const long long per = (n + frame_groups - 1) / frame_groups;
  const long long f_begin = blockIdx.y * per;
  const long long f_end = f_begin + per < n ? f_begin + per : n;
  for (long long t0 = f_begin; t0 < f_end; t0 += WS_FRAMES) {
    const int nf = (int)((f_end - t0) < WS_FRAMES ? (f_end - t0) : WS_FRAMES);
    __syncthreads();
    for (int i = threadIdx.x; i < nf * M * 7; i += blockDim.x) {
      const int f = i / (M * 7), r = i % (M * 7);
      const int m = r / 7, c = r % 7;
      sp[i] = preds[((long long)m * n + t0 + f) * 7 + c];
    }
    for (int i = threadIdx.x; i < nf; i += blockDim.x) sgt[i] = gt[t0 + i];
    __syncthreads();
    if (active) {
      for (int f = 0; f < nf; ++f) {
        const double* pf = sp + f * M * 7;
        double best = 0.0;
        int bi = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          double v = __dmul_rn(pf[c], w[0][c]);
#pragma unroll
          for (int m = 1; m < M; ++m) v = __dadd_rn(v, __dmul_rn(pf[m * 7 + c], w[m][c]));
          if (c == 0) { best = v; bi = 0; }
          else if (v > best || (v != v && best == best)) { best = v; bi = c; }
        }
        const int g = sgt[f];
        if (g >= 0 && g < 7) my[g * 7 + bi] += 1u;
      }
    }
  }
  (void)lane; (void)warp;
  if (active)
    for (int i = 0; i < 49; ++i)
      if (my[i]) atomicAdd(cm + ws * 49 + i, (unsigned long long)my[i]);
}

// Weighted fusion + arg-max over the 7 basic emotions (get_pred_av.py:34-40, get_metrics): final = P_0 * W1[0] * W2[0];
// final += P_m * W1[m] * W2[m]; np.argmax(final, axis=-1) -- float64, left to right, first maximum, NaN counts as largest.
__global__ void fused_argmax_kernel(const double* __restrict__ preds, int n_models, long long n, const double* __restrict__ w1,
                                    const double* __restrict__ w2, int* __restrict__ labels) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  double best = 0.0;
  int bi = 0;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    double v = __dmul_rn(__dmul_rn(preds[f * 7 + c], w1[c]), w2[0]);
    for (int m = 1; m < n_models; ++m)
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(preds[((long long)m * n + f) * 7 + c], w1[m * 7 + c]), w2[m]));
    if (c == 0) { best = v; bi = 0; }
    else if (v > best || (v != v && best == best)) { best = v; bi = c; }
  }
  labels[f] = bi;
}

}  // namespace avcer

using namespace avcer;

template <typename TIn>
static int fuse_compound_impl(const TIn* p_vs, const TIn* p_vd, const TIn* p_a, int64_t n, const double* w1_host,
                              const double* w2_host, int ce_weights_type, int ce_mask, int64_t* labels,
                              int64_t label_pitch, void* stream) {
  AVCER_REQUIRE(n >= 0, "fuse_compound: negative n");
  AVCER_REQUIRE(label_pitch == 0 || label_pitch >= n, "fuse_compound: label_pitch %lld < n %lld", (long long)label_pitch, (long long)n);
  if (n == 0) return 0;
  const long long ld = label_pitch > 0 ? label_pitch : n;
  FuseParams p{};
  p.has_w1 = w1_host != nullptr;
  p.ce_mask = ce_mask != 0;
  if (p.has_w1) {
    AVCER_REQUIRE(w2_host != nullptr, "fuse_compound: weights_2 required with weights_1");
    for (int m = 0; m < 3; ++m) {
      for (int c = 0; c < 7; ++c) p.w[m][c] = w1_host[m * 7 + c];
      p.w2[m] = w2_host[m];
    }
  }
  // Rule-2 weights d[i]/(d[i1]+d[i2]) with d = {1:5,2:6,3:5,4:6,5:4,6:2} (run.py:116-123, utils.py:228-231)
  static const double dict_w[7] = {0, 5, 6, 5, 6, 4, 2};
  static const int pairs[7][2] = {{3, 6}, {4, 6}, {5, 6}, {2, 6}, {1, 6}, {3, 5}, {1, 5}};
  for (int k = 0; k < 7; ++k) {
    if (ce_weights_type) {
      const double s = dict_w[pairs[k][0]] + dict_w[pairs[k][1]];
      p.ce_w[k][0] = dict_w[pairs[k][0]] / s;
      p.ce_w[k][1] = dict_w[pairs[k][1]] / s;
    } else {
      p.ce_w[k][0] = 1.0;
      p.ce_w[k][1] = 1.0;
    }
  }
  p.w2_one = p.has_w1 && p.w2[0] == 1.0 && p.w2[1] == 1.0 && p.w2[2] == 1.0;
  p.cew_one = !ce_weights_type;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p_vs) | reinterpret_cast<uintptr_t>(p_vd) | reinterpret_cast<uintptr_t>(p_a) |
                         reinterpret_cast<uintptr_t>(labels)) & 15) == 0 && ld % 2 == 0;
  constexpr int kFpt = 16 / sizeof(TIn);
  const long long work = aligned ? n / kFpt : n;
  const long long blocks_needed = (work + FUSE_THREADS - 1) / FUSE_THREADS + 1;
  (void)kFpt;
  const long long cap = 1ll << 20;           // one warp-chunk per warp: no second, mostly empty round
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  cudaStream_t st = as_stream(stream);
  // numpy promotes to float64 as soon as the (Python-float) weight lists take part; without
  // them the arithmetic stays in the dtype of the probability arrays.
  const bool f64 = p.has_w1 || sizeof(TIn) == 8;
  long long* lab = reinterpret_cast<long long*>(labels);
  if (aligned) {
    if (f64) fuse_compound_kernel<TIn, double><<<grid, FUSE_THREADS, 0, st>>>(p_vs, p_vd, p_a, n, p, lab, ld);
    else fuse_compound_kernel<TIn, float><<<grid, FUSE_THREADS, 0, st>>>(p_vs, p_vd, p_a, n, p, lab, ld);
  } else {
    if (f64) fuse_compound_scalar_kernel<TIn, double><<<grid, FUSE_THREADS, 0, st>>>(p_vs, p_vd, p_a, n, p, lab, ld);
    else fuse_compound_scalar_kernel<TIn, float><<<grid, FUSE_THREADS, 0, st>>>(p_vs, p_vd, p_a, n, p, lab, ld);
  }
  return check_launch("fuse_compound_kernel");
}

extern "C" int avcer_fuse_compound(const float* p_vs, const float* p_vd, const float* p_a, int64_t n,
                                   const double* w1_host, const double* w2_host, int ce_weights_type, int ce_mask,
                                   int64_t* labels, int64_t label_pitch, void* stream) {
  return fuse_compound_impl<float>(p_vs, p_vd, p_a, n, w1_host, w2_host, ce_weights_type, ce_mask, labels, label_pitch, stream);
}

extern "C" int avcer_fuse_compound_f64(const double* p_vs, const double* p_vd, const double* p_a, int64_t n,
                                       const double* w1_host, const double* w2_host, int ce_weights_type,
                                       int ce_mask, int64_t* labels, int64_t label_pitch, void* stream) {
  return fuse_compound_impl<double>(p_vs, p_vd, p_a, n, w1_host, w2_host, ce_weights_type, ce_mask, labels, label_pitch, stream);
}

extern "C" int avcer_softmax7_f64(const double* x, int64_t n, int ld, double* y, void* stream) {
  AVCER_REQUIRE(n >= 0 && ld >= 7, "softmax7: bad shape");
  if (n == 0) return 0;
  softmax7_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, n, ld, y);
  return check_launch("softmax7_kernel");
}

extern "C" int avcer_softmax7(const float* x, int64_t n, int ld, float* y, void* stream) {
  AVCER_REQUIRE(n >= 0 && ld >= 7, "softmax7: bad shape");
  if (n == 0) return 0;
  softmax7_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, n, ld, y);
  return check_launch("softmax7_kernel");
}

extern "C" int avcer_window_to_frame_mean(const float* logits, int n_win, int ncls, const int32_t* f_lo,
                                          const int32_t* f_hi, int64_t n_frames, float* out, void* stream) {
  AVCER_REQUIRE(n_win >= 0 && ncls > 0 && n_frames >= 0, "window_to_frame_mean: bad shape");
  const long long tot = n_frames * ncls;
  if (tot == 0) return 0;
  window_to_frame_mean_kernel<float><<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(logits, n_win, ncls, f_lo,
                                                                                                    f_hi, n_frames, out);
  return check_launch("window_to_frame_mean_kernel");
}

extern "C" int avcer_window_to_frame_mean_f64(const double* logits, int n_win, int ncls, const int32_t* f_lo,
                                              const int32_t* f_hi, int64_t n_frames, double* out, void* stream) {
  AVCER_REQUIRE(n_win >= 0 && ncls > 0 && n_frames >= 0, "window_to_frame_mean: bad shape");
  const long long tot = n_frames * ncls;
  if (tot == 0) return 0;
  window_to_frame_mean_kernel<double><<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(logits, n_win, ncls, f_lo,
                                                                                                     f_hi, n_frames, out);
  return check_launch("window_to_frame_mean_kernel");
}

extern "C" int avcer_gather_rows(const float* src, const int32_t* src_index, int64_t n_out, int ncols,
                                 const int32_t* perm, float* out, void* stream) {
  AVCER_REQUIRE(n_out >= 0 && ncols > 0, "gather_rows: bad shape");
  const long long tot = n_out * ncols;
  if (tot == 0) return 0;
  gather_rows_kernel<float><<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(src, src_index, n_out, ncols, perm,
                                                                                          out);
  return check_launch("gather_rows_kernel");
}

extern "C" int avcer_gather_rows_f64(const double* src, const int32_t* src_index, int64_t n_out, int ncols,
                                     const int32_t* perm, double* out, void* stream) {
  AVCER_REQUIRE(n_out >= 0 && ncols > 0, "gather_rows: bad shape");
  const long long tot = n_out * ncols;
  if (tot == 0) return 0;
  gather_rows_kernel<double><<<(unsigned)((tot + 255) / 256), 256, 0, as_stream(stream)>>>(src, src_index, n_out, ncols, perm,
                                                                                           out);
  return check_launch("gather_rows_kernel");
}

extern "C" int avcer_compound_scores(const void* pred, int64_t n, int ncols, int pred_f64, const int32_t* pairs_host,
                                     const double* w_host, int k, int ce_mask, double* out, void* stream) {
  AVCER_REQUIRE(k >= 1 && k <= 16 && ncols >= 1 && n >= 0, "compound_scores: bad shape");
  if (n == 0) return 0;
  PairTable t{};
  t.k = k;
  for (int i = 0; i < k; ++i) {
    t.i1[i] = pairs_host[2 * i]; t.i2[i] = pairs_host[2 * i + 1];
    AVCER_REQUIRE(t.i1[i] >= 0 && t.i1[i] < ncols && t.i2[i] >= 0 && t.i2[i] < ncols, "compound_scores: pair index out of range");
    t.w1[i] = w_host[2 * i]; t.w2[i] = w_host[2 * i + 1];
  }
  const long long tot = n * k;
  const unsigned g = (unsigned)((tot + 255) / 256);
  if (pred_f64) compound_scores_kernel<double><<<g, 256, 0, as_stream(stream)>>>((const double*)pred, n, ncols, t, ce_mask, out);
  else compound_scores_kernel<float><<<g, 256, 0, as_stream(stream)>>>((const float*)pred, n, ncols, t, ce_mask, out);
  return check_launch("compound_scores_kernel");
}

extern "C" int avcer_weight_search_confusion(const double* preds, int n_models, int64_t n, const int32_t* gt,
                                             const double* weights, int64_t n_weights, uint64_t* cm, void* stream) {
  AVCER_REQUIRE(n_models == 2 || n_models == 3, "weight_search: 2 or 3 prediction streams");
  AVCER_REQUIRE(n >= 0 && n_weights >= 0, "weight_search: bad sizes");
  if (n == 0 || n_weights == 0) return 0;
  const int wblocks = (int)((n_weights + WS_WARPS * 32 - 1) / (WS_WARPS * 32));
  long long tiles = (n + WS_FRAMES - 1) / WS_FRAMES;
  int fg = (int)(tiles < 1 ? 1 : tiles);
  const int want = (148 * 2 + wblocks - 1) / wblocks;          // enough CTAs to fill the GPU twice
  if (fg > want) fg = want;
  if (fg > 65535) fg = 65535;
  const size_t smem = (size_t)WS_FRAMES * n_models * 7 * sizeof(double) + WS_FRAMES * sizeof(int) +
                      (size_t)WS_WARPS * 32 * 49 * sizeof(unsigned);
  dim3 grid(wblocks, fg);
  cudaStream_t st = as_stream(stream);
  if (n_models == 3) {
    AVCER_CUDA(cudaFuncSetAttribute(weight_search_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weight_search_kernel<3><<<grid, WS_WARPS * 32, smem, st>>>(preds, n, gt, weights, n_weights, fg, (unsigned long long*)cm);
  } else {
    AVCER_CUDA(cudaFuncSetAttribute(weight_search_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weight_search_kernel<2><<<grid, WS_WARPS * 32, smem, st>>>(preds, n, gt, weights, n_weights, fg, (unsigned long long*)cm);
  }
  return check_launch("weight_search_kernel");
}

extern "C" int avcer_fused_argmax(const double* preds, int n_models, int64_t n, const double* w1, const double* w2,
                                  int32_t* labels, void* stream) {
  AVCER_REQUIRE(n_models >= 1 && n >= 0, "fused_argmax: bad sizes");
  if (n == 0) return 0;
  fused_argmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(preds, n_models, (long long)n, w1, w2, labels);
  return check_launch("fused_argmax_kernel");
}
