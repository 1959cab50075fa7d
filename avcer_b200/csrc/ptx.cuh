// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is device-only glue; no algorithmic content.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace avcer {

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), the "gelu" of HF wav2vec2 (erf form, not the tanh form).
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// Same function with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. far below one
// bf16 ulp): one MUFU.RCP + one MUFU.EX2 + a 5-term Horner instead of libdevice erff.  Used where
// the result is stored as bf16; the fp32 mode keeps erff.
// 1 / d for d >= 1 (no denormal operand or result to rescue: the bare MUFU.RCP, without the range fix-up
// FSETP / FSEL / 2 x FMUL that __fdividef and 1.0f / d carry).
__device__ __forceinline__ float rcp_approx(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);        // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// (a0, a1) += (b0, b1) as ONE packed fp32x2 instruction (sm_100 FADD2; each lane rounds like add.rn.f32).  Epilogues
// are bound by instruction issue on the four schedulers, so halving the add count is a direct win.
__device__ __forceinline__ void add_f32x2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}

// ---------------------------------------------------------------- packed fp32x2 arithmetic (sm_100 FMUL2 / FFMA2)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// gelu_erf_fast on two values at once: the same Abramowitz & Stegun 7.1.26 evaluation with every multiply / FMA
// issued as a packed fp32x2 instruction (10 packed + 4 MUFU + 2 logic per PAIR instead of ~17 per value).  The layers
// that apply GELU to every activation (conv0, the feature-extractor LayerNorms, the FFN epilogue) are issue-bound.
// With a = |x|, z = a / sqrt 2:  t = 1 / (1 + p z),  erf(z) = 1 - t * poly(t) * exp(-z^2),  gelu = 0.5 (x + a erf(z));
// the constants of z are folded into the coefficients and the polynomial is evaluated negated, so no instruction is
// spent on z itself, on a sign flip or on halving both terms.
__device__ __forceinline__ uint64_t gelu_erf_fast2(uint64_t x) {
  const uint64_t a = x & 0x7fffffff7fffffffull;
  const uint64_t d = fma_f32x2(pack_f32x2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f), a,
                               pack_f32x2(1.0f, 1.0f));
  float d0, d1;
  unpack_f32x2(d, d0, d1);
  const uint64_t t = pack_f32x2(rcp_approx(d0), rcp_approx(d1));
  uint64_t poly = fma_f32x2(pack_f32x2(-1.061405429f, -1.061405429f), t, pack_f32x2(1.453152027f, 1.453152027f));
  poly = fma_f32x2(poly, t, pack_f32x2(-1.421413741f, -1.421413741f));
  poly = fma_f32x2(poly, t, pack_f32x2(0.284496736f, 0.284496736f));
  poly = fma_f32x2(poly, t, pack_f32x2(-0.254829592f, -0.254829592f));
  const uint64_t npt = mul_f32x2(poly, t);                // -t * poly(t)
  // exp(-z^2) = 2^(-a^2 * log2(e) / 2)
  const uint64_t nz = mul_f32x2(mul_f32x2(a, pack_f32x2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f)), a);
  float n0, n1;
  unpack_f32x2(nz, n0, n1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(n0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(n1));
  const uint64_t er = fma_f32x2(npt, pack_f32x2(e0, e1), pack_f32x2(1.0f, 1.0f));      // erf(|x| / sqrt 2)
  return mul_f32x2(fma_f32x2(a, er, x), pack_f32x2(0.5f, 0.5f));
}
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
  unpack_f32x2(gelu_erf_fast2(pack_f32x2(x0, x1)), x0, x1);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
#ifndef AVCER_MBAR_TIMEOUT_NS
#define AVCER_MBAR_TIMEOUT_NS 4000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > AVCER_MBAR_TIMEOUT_NS) __trap();
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// 1-D bulk copy global -> shared (contiguous bytes, multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// TMA store (shared -> global, tiled, bulk async-group completion); partial boxes are clipped.
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint4& v) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 operand tile written by TMA with
// SWIZZLE_128B (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart) or SWIZZLE_64B
// (rows of 32 bf16 = 64 B, 8-row groups 512 B apart). Field layout: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t sbo_bytes,
                                                     uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                    // LBO: unused for swizzled K-major
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;          // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
  return d;
}

// SWIZZLE_NONE K-major descriptor with explicit leading (K-chunk) and stride (8-row group) byte offsets.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  // bits 4-5: accumulator format (1 = f32); bits 7-9 / 10-12: A / B element format (1 = bf16, 0 = f16: the AVCER_HALF build)
#ifdef AVCER_HALF
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
#else
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
#endif
}

}  // namespace avcer
