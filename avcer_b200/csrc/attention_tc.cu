// Tensor-core multi-head self-attention for short sequences (T <= 208: one 4 s audio window is
// T = 199 tokens), bf16 in / fp32 accumulate / bf16 out.  Replaces the eager
// softmax(Q K^T / sqrt(d)) V of HF Wav2Vec2Attention (12 encoder layers, 16 heads x 64) and of
// src/architectures/attention_layers.py:10-38 (tl1: 32 heads x 32, tl2: 16 heads x 64).
//
// One CTA of 7 warps per (window, head); two CTAs share an SM, so the staging of one overlaps the math of the
// other.  K and V of the head are staged once in shared memory; every warp owns up to two 16-row query tiles.
// The key axis is walked ONCE in two blocks of <= 112 keys whose scores stay in registers (online softmax:
// block maximum, rescale of the running output between the blocks, exp2, row sums, PV), so QK^T is computed a
// single time; QK^T and PV run on the tensor cores (mma.sync m16n8k16, the probabilities are re-used directly as
// the A operand of PV).  This op is ~2 % of the audio network's FLOPs; the dense projections around it are tcgen05.
#include "common.h"

namespace avcer {

constexpr int ATT_MAXT = 208;               // 13 x 16
constexpr int ATT_WARPS = 7;                // 13 query tiles over 7 warps (two tiles per warp)
constexpr int ATT_BLK = 7;                  // key chunks (of 16) per softmax block: 2 blocks cover 208 keys

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef AVCER_HALF
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32, 2)
attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv, int t, int heads, float scale_log2e,
                    __nv_bfloat16* __restrict__ out) {
  constexpr int PITCH = DH + 8;                       // +16 B per row: conflict-free ldmatrix
  constexpr int KS = DH / 16;                         // k-steps of QK^T
  constexpr int NT = DH / 8;                          // n-tiles of PV
  extern __shared__ __align__(16) unsigned char att_smem[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sk = sq + ATT_MAXT * PITCH;
  __nv_bfloat16* sv = sk + ATT_MAXT * PITCH;
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row_stride = 3ll * heads * DH;
  const __nv_bfloat16* base = qkv + (long long)b * t * row_stride + (long long)head * DH;

  pdl_wait();
  pdl_launch_dependents();
  // stage Q, K, V rows (16 B vectors), zero-fill the padded rows
  constexpr int VEC = DH / 8;
  for (int i = tid; i < ATT_MAXT * VEC * 3; i += blockDim.x) {
    const int which = i / (ATT_MAXT * VEC);
    const int rem = i % (ATT_MAXT * VEC);
    const int r = rem / VEC, v = rem % VEC;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < t) val = __ldg(reinterpret_cast<const uint4*>(base + r * row_stride + (long long)which * heads * DH) + v);
    __nv_bfloat16* dst = (which == 0 ? sq : which == 1 ? sk : sv) + r * PITCH + v * 8;
    *reinterpret_cast<uint4*>(dst) = val;
  }
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  const int n_chunks = (t + 15) / 16;
  __nv_bfloat16* ob = out + (long long)b * t * heads * DH + (long long)head * DH;

  for (int q0 = warp * 16; q0 < t; q0 += ATT_WARPS * 16) {
    // Q fragments of this warp's 16 rows
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int r = q0 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int c = ks * 16 + 8 * (lane >> 4);
      ldsm_x4(smem_u32_generic(sq + r * PITCH + c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows g and g+8 of the warp tile

#pragma unroll 1
    for (int c0 = 0; c0 < n_chunks; c0 += ATT_BLK) {
      // scores of up to ATT_BLK key chunks: s[c][0..3] = keys 16ch..+7, s[c][4..7] = keys +8..+15
      float s[ATT_BLK][8];
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATT_BLK; ++c) {
        const int ch = c0 + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) s[c][i] = (ch < n_chunks) ? 0.f : -INFINITY;
        if (ch < n_chunks) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t b0, b1, b2, b3;
            const int key = ch * 16 + (lane & 7) + 8 * (lane >> 4);
            const int d = ks * 16 + 8 * ((lane >> 3) & 1);
            ldsm_x4(smem_u32_generic(sk + key * PITCH + d), b0, b1, b2, b3);
            mma_bf16(*reinterpret_cast<float(*)[4]>(&s[c][0]), qf[ks], b0, b1);
            mma_bf16(*reinterpret_cast<float(*)[4]>(&s[c][4]), qf[ks], b2, b3);
          }
          // mask keys >= t
          const int k0 = ch * 16 + 2 * tq;
          if (k0 >= t) { s[c][0] = -INFINITY; s[c][2] = -INFINITY; }
          if (k0 + 1 >= t) { s[c][1] = -INFINITY; s[c][3] = -INFINITY; }
          if (k0 + 8 >= t) { s[c][4] = -INFINITY; s[c][6] = -INFINITY; }
          if (k0 + 9 >= t) { s[c][5] = -INFINITY; s[c][7] = -INFINITY; }
          bm0 = fmaxf(bm0, fmaxf(fmaxf(s[c][0], s[c][1]), fmaxf(s[c][4], s[c][5])));
          bm1 = fmaxf(bm1, fmaxf(fmaxf(s[c][2], s[c][3]), fmaxf(s[c][6], s[c][7])));
        }
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      // online softmax: every chunk of a block starts below t, so the block maximum is finite
      const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
      if (c0 > 0) {
        const float a0 = exp2f((m0 - n0) * scale_log2e), a1 = exp2f((m1 - n1) * scale_log2e);
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int n = 0; n < NT; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
      }
      m0 = n0; m1 = n1;
      const float mb0 = m0 * scale_log2e, mb1 = m1 * scale_log2e;
#pragma unroll
      for (int c = 0; c < ATT_BLK; ++c) {
        const int ch = c0 + c;
        if (ch < n_chunks) {
          float p[8];
          p[0] = exp2f(s[c][0] * scale_log2e - mb0); p[1] = exp2f(s[c][1] * scale_log2e - mb0);
          p[2] = exp2f(s[c][2] * scale_log2e - mb1); p[3] = exp2f(s[c][3] * scale_log2e - mb1);
          p[4] = exp2f(s[c][4] * scale_log2e - mb0); p[5] = exp2f(s[c][5] * scale_log2e - mb0);
          p[6] = exp2f(s[c][6] * scale_log2e - mb1); p[7] = exp2f(s[c][7] * scale_log2e - mb1);
          l0 += p[0] + p[1] + p[4] + p[5];
          l1 += p[2] + p[3] + p[6] + p[7];
          uint32_t pa[4] = {pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])};
#pragma unroll
          for (int n2 = 0; n2 < NT / 2; ++n2) {
            uint32_t v0, v1, v2, v3;
            const int key = ch * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
            const int d = n2 * 16 + 8 * (lane >> 4);
            ldsm_x4_t(smem_u32_generic(sv + key * PITCH + d), v0, v1, v2, v3);
            mma_bf16(o[2 * n2], pa, v0, v1);
            mma_bf16(o[2 * n2 + 1], pa, v2, v3);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int c = n * 8 + 2 * tq;
      if (r0 < t) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * heads * DH + c) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
      if (r1 < t) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * heads * DH + c) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
    }
  }
}

template <int DH>
static int launch_attention_tc(const void* qkv, int n, int t, int heads, float scale, void* out, cudaStream_t st) {
  constexpr int PITCH = DH + 8;
  const size_t smem = (size_t)3 * ATT_MAXT * PITCH * sizeof(__nv_bfloat16);
  auto kern = attention_tc_kernel<DH>;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(heads, n);
  launch_pdl(kern, grid, ATT_WARPS * 32, smem, st, static_cast<const __nv_bfloat16*>(qkv), t, heads,
             scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out));
  return check_launch("attention_tc_kernel");
}

int attention_tc5(const void* qkv, int n, int t, int heads, float scale, void* out, cudaStream_t st);   // contract.cu (tcgen05)

int attention_tc(const void* qkv, int n, int t, int heads, int dh, float scale, void* out, cudaStream_t st) {
  static const int att5 = getenv("AVCER_ATT5") ? atoi(getenv("AVCER_ATT5")) : 1;
  if (att5 != 0 && dh == 64 && t <= ATT_MAXT) return attention_tc5(qkv, n, t, heads, scale, out, st);
  AVCER_REQUIRE(t >= 1 && t <= ATT_MAXT, "attention(bf16): T=%d exceeds the on-chip limit %d", t, ATT_MAXT);
  AVCER_REQUIRE(n <= 65535, "attention: at most 65535 windows per call");
  if (dh == 64) return launch_attention_tc<64>(qkv, n, t, heads, scale, out, st);
  if (dh == 32) return launch_attention_tc<32>(qkv, n, t, heads, scale, out, st);
  return set_error("attention: head dim must be 32 or 64");
}

}  // namespace avcer
