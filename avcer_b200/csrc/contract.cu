// avcer_contract: host launcher of the tcgen05 implicit-GEMM kernel (bf16) and the SIMT fp32
// kernel ("fp32 mode"), both driven by the same avcer_contract_desc geometry.
#include "common.h"
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "conv3x3.cuh"
#include "stem_pool.cuh"
#include "attention_tc5.cuh"
#include "conv0_tc.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>

namespace avcer {

// ------------------------------------------------------------------ tensor-map encoding
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box,
                      CUtensorMapSwizzle swz, const uint32_t* elem_strides = nullptr) {
  auto fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
#ifdef AVCER_HALF
  constexpr CUtensorMapDataType kElem = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#else
  constexpr CUtensorMapDataType kElem = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#endif
  CUresult r = fn(m, kElem, rank, const_cast<void*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] "
        "stride1 %llu base %p",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
        rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, (unsigned long long)strides_bytes[0], base);
  }
  return 0;
}

// Pick the box (bw, bh, bn), bw*bh*bn <= 128, that covers W x H x NB with the fewest M tiles.
static void choose_box(int W, int H, int NB, int* bw, int* bh, int* bn) {
  long long best = -1;
  int rb = 0;
  for (int w = 1; w <= 128 && w <= W; ++w) {
    for (int h = 1; h * w <= 128 && h <= H; ++h) {
      int n = 128 / (w * h);
      if (n > NB) n = NB;
      if (n > 256) n = 256;
      const long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((NB + n - 1) / n);
      const int r = w * h * n;
      // fewer tiles first; then fuller boxes of wider rows (better locality of the TMA box)
      if (best < 0 || tiles < best || (tiles == best && (w > *bw || (w == *bw && r > rb)))) {
        best = tiles; *bw = w; *bh = h; *bn = n; rb = r;
      }
    }
  }
}

template <int BN, int BK, int MODE, int KSUB, int OCC>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                     const TcGemmParams& p, int grid, cudaStream_t st) {
  using Cfg = TcGemmCfg<BN, BK, MODE, KSUB, OCC>;
  auto kern = tc_gemm_kernel<BN, BK, MODE, KSUB, OCC>;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  int g = grid * OCC;                     // persistent: OCC CTAs per SM
  if (g > p.num_tiles) g = p.num_tiles;
  launch_pdl_tpc(kern, g, Cfg::THREADS, Cfg::SMEM, st, ta, tb, tc, tr, p);
  return check_launch("tc_gemm_kernel");
}

template <int MODE, int BN, int RS = 4, int FLAT = 0>
static int launch_tc2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                      const TcGemmParams& p, int sms, cudaStream_t st) {
  using Cfg = TcGemm2Cfg<MODE, BN, RS, FLAT>;
  auto kern = tc_gemm2_kernel<MODE, BN, RS, FLAT>;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  const int m_tiles = p.tw * p.th * p.tn;
  const int pair_tiles = ((m_tiles + 1) / 2) * p.tiles_n;
  int clusters = sms / 2;
  if (clusters > pair_tiles) clusters = pair_tiles;
  launch_pdl(kern, 2 * clusters, Cfg::THREADS, Cfg::SMEM, st, ta, tb, tc, tr, p);     // __cluster_dims__(2,1,1)
  return check_launch("tc_gemm2_kernel");
}

// Development aid (avcer_debug_set_trace in the header): device buffer of kTraceCtas*kTraceTiles*kTraceSlots u64 that the
// two-SM kernel fills with per-tile clock64 stamps of its first CTAs; nullptr (default) disables tracing.
static unsigned long long* g_trace = nullptr;
extern "C" int avcer_debug_set_trace(void* buf) {
  g_trace = static_cast<unsigned long long*>(buf);
  return 0;
}

static int num_sms_cached() { return num_sms(); }

// Which kernel the last avcer_contract call of this thread dispatched to (avcer_last_contract_kernel): lets the host
// attribute per-launch timings to the exact kernel (bench.py's roofline names ONE dominant kernel, not a family).
static thread_local int g_last_kernel = 0;
enum KernelId : int {
  KID_NONE = 0, KID_SIMT = 1, KID_TC_64 = 2, KID_TC_128 = 3, KID_TC_256 = 4, KID_TC_F32OUT = 5, KID_TC2_256 = 6, KID_TC2_256_RES = 7,
  KID_TC2_256_FLAT = 8, KID_TC2_256_FLAT_RES = 9, KID_TC2_128 = 10, KID_CONV3_64 = 11, KID_CONV3_128 = 12, KID_TC_STRIP = 13
};

// 3x3 "same" stride-1 convolutions with 64 / 128 output channels over dense NHWC tensors (ResNet-50 layer1 / layer2
// conv2): halo-in-shared-memory kernel of conv3x3.cuh.  Returns -1 when the geometry is not its case.
template <int BN, bool RES>
static int launch_conv3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, Conv3Params& p, cudaStream_t st) {
  using Cfg = Conv3Cfg<BN, RES>;
  auto kern = conv3x3_kernel<BN, RES>;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BUDGET));
    attr_done = true;
  }
  p.a_stages = (Cfg::BUDGET - 1024 - Cfg::B_BYTES - Cfg::C_BYTES) / (int)p.a_stage;
  if (p.a_stages > Cfg::MAX_A_STAGES) p.a_stages = Cfg::MAX_A_STAGES;
  AVCER_REQUIRE(p.a_stages >= 2, "conv3x3: activation stage of %u bytes does not fit twice", p.a_stage);
  const size_t smem = (size_t)p.a_stages * p.a_stage + Cfg::B_BYTES + Cfg::C_BYTES + 1024;
  int g = num_sms();
  if (g > p.num_tiles) g = p.num_tiles;
  launch_pdl_tpc(kern, g, Cfg::THREADS, smem, st, ta, tb, tc, p);
  return check_launch("conv3x3_kernel");
}

static int conv3x3_tc(const avcer_contract_desc* d, cudaStream_t st) {
  static const int on = getenv("AVCER_CONV3") ? atoi(getenv("AVCER_CONV3")) : 1;
  const int C = d->cin, W = d->W, H = d->H, NB = d->NB, Cout = d->cout;
  const bool shape = on != 0 && d->taps_w == 3 && d->taps_h == 3 && d->off_w == -1 && d->off_h == -1 && !d->tap_h_in_dim4 &&
                     d->group_cin_shift == 0 && !d->a_strip && d->a_step <= 1 && !d->out_f32 && d->residual == nullptr && C % 64 == 0 &&
                     ((Cout == 64 && C == 64) || Cout == 128) && W + 2 >= 28 && W + 2 <= 63 && NB >= 1 &&
                     (d->act == ACT_NONE || d->act == ACT_RELU);
  const bool dense = d->a_dim[0] == C && d->a_dim[1] == W && d->a_dim[2] == H && d->a_dim[3] == NB && d->a_stride[0] == 1 &&
                     d->a_stride[1] == C && d->a_stride[2] == (int64_t)W * C && d->a_stride[3] == (int64_t)H * W * C &&
                     d->out_stride[0] >= Cout && d->out_stride[0] % 8 == 0 && d->out_stride[1] == (int64_t)W * d->out_stride[0] &&
                     d->out_stride[2] == (int64_t)H * W * d->out_stride[0];          // output pixels may sit in wider rows
  const bool aligned = ((reinterpret_cast<uintptr_t>(d->a) | reinterpret_cast<uintptr_t>(d->wt) | reinterpret_cast<uintptr_t>(d->out)) & 15) == 0;
  if (!shape || !dense || !aligned) return -1;       // the generic path validates and reports
  Conv3Params p{};
  p.H = H; p.W = W; p.NB = NB; p.C = C; p.Cout = Cout;
  p.P = W + 2;
  p.bh = 256 / p.P;
  if (p.bh > H) p.bh = H;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.num_tiles = NB * p.tiles_h;
  p.kchunks = C / 64;
  p.a_bytes = (unsigned)((p.bh + 2) * p.P) * 128u;
  p.a_stage = (unsigned)(((2 * p.P + 2 + 256) * 128 + 1023) / 1024 * 1024);
  p.bias = d->bias;
  p.act = d->act;
  p.reverse = d->reverse_tiles != 0;
  CUtensorMap ta, tb, tc;
  {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 1};
    uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)1 << 30};
    uint32_t box[5] = {64u, (uint32_t)p.P, (uint32_t)(p.bh + 2), 1u, 1u};
    if (encode_map(&ta, d->a, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)Cout};
    uint64_t strides[1] = {(uint64_t)9 * C * 2};
    uint32_t box[2] = {64u, (uint32_t)Cout};
    if (encode_map(&tb, d->wt, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)NB, 1};
    const uint64_t op = (uint64_t)d->out_stride[0];
    uint64_t strides[4] = {op * 2, (uint64_t)W * op * 2, (uint64_t)H * W * op * 2, (uint64_t)1 << 30};
    uint32_t box[5] = {64u, (uint32_t)W, (uint32_t)p.bh, 1u, 1u};
    if (encode_map(&tc, d->out, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  g_last_kernel = Cout == 64 ? KID_CONV3_64 : KID_CONV3_128;
  if (Cout == 64) return launch_conv3<64, true>(ta, tb, tc, p, st);
  return launch_conv3<128, false>(ta, tb, tc, p, st);
}

static int contract_tc(const avcer_contract_desc* d, cudaStream_t st) {
  {
    const int rc = conv3x3_tc(d, st);
    if (rc >= 0) return rc;
  }
  AVCER_REQUIRE(d->a_stride[0] == 1, "contract: a_stride[0] must be 1");
  const int BK = (d->cin % 64 == 0) ? 64 : 32;
  AVCER_REQUIRE(d->cin % BK == 0, "contract(bf16): cin=%d must be a multiple of 32", d->cin);
  AVCER_REQUIRE(d->cout % 64 == 0, "contract(bf16): cout=%d must be a multiple of 64", d->cout);
  AVCER_REQUIRE(!(d->out_f32 && d->residual), "contract(bf16): fp32 output does not take a residual");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->wt) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                "contract(bf16): a/wt/out must be 16-byte aligned");
  int BN = (d->group_cin_shift != 0 || d->cout % 128 != 0) ? 64 : 128;
  if (BK == 32) BN = 64;   // stem: Cout = 64
  // 128x256 tiles halve the shared-memory operand traffic per MMA (measured 1.45x the MMA rate of 128x128);
  // pick them when the wave-quantised time estimate is lower.
  if (BN == 128 && BK == 64 && d->cout % 256 == 0 && !d->out_f32) {
    const long long mt = ((long long)d->W * d->H * d->NB + 127) / 128;
    const long long sms = num_sms_cached();
    const long long waves128 = (mt * (d->cout / 128) + sms - 1) / sms;
    const long long waves256 = (mt * (d->cout / 256) + sms - 1) / sms;
    if (waves256 * 2.0 / 1.4 < (double)waves128) BN = 256;
  }
  if (getenv("AVCER_NO_BN256") && BN == 256) BN = 128;
  static const int cta2_env = getenv("AVCER_CTA2") ? atoi(getenv("AVCER_CTA2")) : 1;
  const bool use_cta2 = cta2_env != 0 && (BN == 256 || (BN == 128 && cta2_env >= 2)) && BK == 64 && !d->a_strip &&
                        d->group_cin_shift == 0 && !d->out_f32;

  TcGemmParams p{};
  if (d->a_strip) {
    const int64_t strip_elems = d->a_dim[0] * d->a_dim[1];
    AVCER_REQUIRE(d->a_dim[0] % 8 == 0 && d->a_dim[0] <= 64 && d->W <= 128 && d->taps_w == 1 && d->tap_h_in_dim4 &&
                      d->cin == BK && strip_elems >= 8 * (int64_t)(d->W - 1) + BK && d->a_dim[1] <= 256 &&
                      d->a_stride[1] == d->a_dim[0] && d->group_cin_shift == 0 && d->wt_packed != nullptr,
                  "contract(bf16): bad strip geometry");
    p.bw = d->W; p.bh = 1; p.bn = 1;
    p.a_strip = 1;
    p.a_bytes = (unsigned)(strip_elems * 2);
    p.b_packed = d->wt_packed;
    AVCER_REQUIRE(p.a_bytes <= 128u * BK * 2u, "contract(bf16): strip larger than the A stage");
    AVCER_REQUIRE(16 * 127 + BK * 2 <= 128 * BK * 2, "contract(bf16): strip exceeds the A stage");
  } else {
    choose_box(d->W, d->H, d->NB, &p.bw, &p.bh, &p.bn);
    p.a_bytes = (unsigned)(p.bw * p.bh * p.bn) * BK * 2;
  }
  p.tw = (d->W + p.bw - 1) / p.bw;
  p.th = (d->H + p.bh - 1) / p.bh;
  p.tn = (d->NB + p.bn - 1) / p.bn;
  p.tiles_n = (d->cout + BN - 1) / BN;
  const long long nt = (long long)p.tw * p.th * p.tn * p.tiles_n;
  AVCER_REQUIRE(nt < (1ll << 31), "contract: too many tiles");
  p.num_tiles = (int)nt;
  p.W = d->W; p.H = d->H; p.NB = d->NB;
  p.taps_w = d->taps_w; p.taps_h = d->taps_h; p.off_w = d->off_w; p.off_h = d->off_h;
  p.tap_h_in_dim4 = d->tap_h_in_dim4;
  const int a_step = d->a_step > 1 ? d->a_step : 1;
  AVCER_REQUIRE(a_step == 1 || (a_step <= 8 && !d->a_strip && !d->tap_h_in_dim4 && p.bw * a_step <= 256 && p.bh * a_step <= 256),
                "contract(bf16): a_step=%d needs plain (c, w, h, n) input dims and boxes of at most 256 / a_step pixels", a_step);
  p.a_step = a_step;
  p.kchunks = d->cin / BK;
  p.a_c0_per_ntile = d->group_cin_shift;   // BN == 64 == one group per N tile
  p.Cout = d->cout;
  p.out_sw = d->out_stride[0]; p.out_sh = d->out_stride[1]; p.out_sn = d->out_stride[2];
  p.bias = d->bias;
  p.out = d->out;
  p.act = d->act;
  p.res_after_act = d->res_after_act;
  p.reverse = d->reverse_tiles != 0;
  p.trace = g_trace;
  if (p.num_tiles == 0) return 0;

  const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ta, tb;
  {
    uint64_t dims[5], strides[4];
    for (int i = 0; i < 5; ++i) dims[i] = (uint64_t)d->a_dim[i];
    for (int i = 1; i < 5; ++i) {
      AVCER_REQUIRE(d->a_stride[i] % 8 == 0, "contract(bf16): a_stride[%d]=%lld must be a multiple of 8 elements",
                    i, (long long)d->a_stride[i]);
      strides[i - 1] = (uint64_t)d->a_stride[i] * 2;
    }
    uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, 1u};
    if (d->a_strip) { box[0] = (uint32_t)d->a_dim[0]; box[1] = (uint32_t)d->a_dim[1]; box[2] = 1; box[3] = 1; }
    // strided spatial conv: the box spans bw*s x bh*s input pixels and TMA keeps every s-th one (traversal stride)
    uint32_t estr[5] = {1u, (uint32_t)a_step, (uint32_t)a_step, 1u, 1u};
    if (a_step > 1) { box[1] *= (uint32_t)a_step; box[2] *= (uint32_t)a_step; }
    if (encode_map(&ta, d->a, 5, dims, strides, box, d->a_strip ? CU_TENSOR_MAP_SWIZZLE_NONE : swz, a_step > 1 ? estr : nullptr)) return 1;
  }
  {
    const uint64_t ktot = (uint64_t)d->taps_w * d->taps_h * d->cin;
    uint64_t dims[2] = {ktot, (uint64_t)d->cout};
    uint64_t strides[1] = {ktot * 2};
    AVCER_REQUIRE((ktot * 2) % 16 == 0, "contract: weight row pitch must be a multiple of 16 bytes");
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_cta2 ? BN / 2 : BN)};   // two-SM tiles: each CTA stages half of the 256 weight rows
    if (encode_map(&tb, d->wt, 2, dims, strides, box, swz)) return 1;
  }
  // output / residual maps: (cout, w, h, n) boxes of the same shape as the activation box
  CUtensorMap tc = ta, tr = ta;
  const int mode = d->out_f32 ? OUT_DIRECT_F32 : (d->residual ? OUT_TMA_RES : OUT_TMA);
  // Plain [M, Cout] outputs (Linear layers, pointwise convs run as GEMMs): M tiles are 128 consecutive rows, so the
  // two-SM kernel can use its barrier-free per-warp epilogue.
  static const int flat_env = getenv("AVCER_FLAT") ? atoi(getenv("AVCER_FLAT")) : 1;
  // It trades two pipeline stages for per-warp staging slabs: only for short K loops, whose tiles are epilogue-bound
  // (deep K loops are MMA-bound and want the stages: measured +30 % time on K >= 1024 GEMMs with residual).
  const bool flat = flat_env != 0 && use_cta2 && BN == 256 && d->H == 1 && d->NB == 1 && p.bw == 128 && p.bh == 1 && p.bn == 1 &&
                    (long long)d->taps_w * d->taps_h * p.kchunks <= (flat_env > 1 ? 1 << 30 : 8);
  if (mode != OUT_DIRECT_F32) {
    const int64_t ext[3] = {d->W, d->H, d->NB};
    auto make_out_map = [&](CUtensorMap* m, const void* base, const int64_t* str) -> int {
      uint64_t dims[5] = {(uint64_t)d->cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->NB, 1};
      uint64_t strides[4];
      for (int i = 0; i < 3; ++i) {
        int64_t sv = ext[i] > 1 ? str[i] : (int64_t)1 << 24;          // extent-1 dims: any legal pitch
        if (sv % 8 != 0 || sv <= 0) return set_error("contract(bf16): out/residual stride[%d]=%lld must be a positive multiple of 8", i, (long long)sv);
        strides[i] = (uint64_t)sv * 2;
      }
      strides[3] = (uint64_t)1 << 30;
      uint32_t box[5] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, 1u};
      if (flat) box[1] = 32u;                                   // per-warp slabs of 32 rows (FLAT epilogue)
      return encode_map(m, base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    if (make_out_map(&tc, d->out, d->out_stride)) return 1;
    if (mode == OUT_TMA_RES) {
      AVCER_REQUIRE((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "contract(bf16): residual must be 16-byte aligned");
      if (make_out_map(&tr, d->residual, d->res_stride)) return 1;
    }
  }
  const int grid = num_sms_cached();
  if (use_cta2)
    g_last_kernel = BN != 256 ? KID_TC2_128 : flat ? (mode == OUT_TMA ? KID_TC2_256_FLAT : KID_TC2_256_FLAT_RES)
                                                   : (mode == OUT_TMA ? KID_TC2_256 : KID_TC2_256_RES);
  else
    g_last_kernel = mode == OUT_DIRECT_F32 ? KID_TC_F32OUT : d->a_strip ? KID_TC_STRIP : BN == 64 ? KID_TC_64 : BN == 128 ? KID_TC_128 : KID_TC_256;
  if (use_cta2) {
    if (BN == 256) {
      if (flat) {
        if (mode == OUT_TMA) return launch_tc2<OUT_TMA, 256, 4, 1>(ta, tb, tc, tr, p, grid, st);
        return launch_tc2<OUT_TMA_RES, 256, 4, 1>(ta, tb, tc, tr, p, grid, st);
      }
      if (mode == OUT_TMA) return launch_tc2<OUT_TMA, 256>(ta, tb, tc, tr, p, grid, st);
      static const int rslots = getenv("AVCER_RSLOTS") ? atoi(getenv("AVCER_RSLOTS")) : 4;
      if (rslots == 2) return launch_tc2<OUT_TMA_RES, 256, 2>(ta, tb, tc, tr, p, grid, st);
      return launch_tc2<OUT_TMA_RES, 256, 4>(ta, tb, tc, tr, p, grid, st);
    }
    if (mode == OUT_TMA) return launch_tc2<OUT_TMA, 128>(ta, tb, tc, tr, p, grid, st);
    return launch_tc2<OUT_TMA_RES, 128>(ta, tb, tc, tr, p, grid, st);
  }
  // K chunks per pipeline stage: small tiles (little MMA work per chunk) batch several chunks per barrier
#define AVCER_TC_CASE(bn, bk, occ)                                                                         \
  if (BN == bn && BK == bk) {                                                                              \
    if (mode == OUT_TMA) return launch_tc<bn, bk, OUT_TMA, 1, occ>(ta, tb, tc, tr, p, grid, st);           \
    if (mode == OUT_TMA_RES) return launch_tc<bn, bk, OUT_TMA_RES, 1, occ>(ta, tb, tc, tr, p, grid, st);   \
    return launch_tc<bn, bk, OUT_DIRECT_F32, 1, 1>(ta, tb, tc, tr, p, grid, st);                           \
  }
  static const int occ128 = getenv("AVCER_OCC128") ? atoi(getenv("AVCER_OCC128")) : 1;
  static const int occ64 = getenv("AVCER_OCC64") ? atoi(getenv("AVCER_OCC64")) : 2;
  AVCER_TC_CASE(256, 64, 1)
  if (occ128 == 2) { AVCER_TC_CASE(128, 64, 2) } else { AVCER_TC_CASE(128, 64, 1) }
  if (occ64 == 2) { AVCER_TC_CASE(64, 64, 2) AVCER_TC_CASE(64, 32, 2) } else { AVCER_TC_CASE(64, 64, 1) AVCER_TC_CASE(64, 32, 1) }
#undef AVCER_TC_CASE
  return set_error("contract: no tensor-core instantiation for BN=%d BK=%d", BN, BK);
}

// ------------------------------------------------------------------ fused stem + max-pool (bf16)
static int stem_pool_tc(const void* x, const void* w_packed, const float* bias, int n, void* out, int64_t out_pitch, cudaStream_t st) {
  AVCER_REQUIRE(out_pitch >= 64 && out_pitch % 8 == 0, "stem_pool: out_pitch %lld must be a multiple of 8, at least 64", (long long)out_pitch);
  AVCER_REQUIRE(x != nullptr && w_packed != nullptr && bias != nullptr && out != nullptr, "stem_pool: null pointer");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "stem_pool: x / w_packed / out must be 16-byte aligned");
  if (n == 0) return 0;
  // strips of the zero-bordered [n, 232, 240, 4] input: (64 elements, 15 per padded row, output row oy -> padded row
  // 2*oy, crop, filter row ky -> +1 padded row)
  constexpr uint64_t ROW = 240 * 4, IMG = 232 * ROW;
  CUtensorMap ta;
  uint64_t dims[5] = {64, ROW / 64, 112, (uint64_t)n, 7};
  uint64_t strides[4] = {64 * 2, 2 * ROW * 2, IMG * 2, ROW * 2};
  uint32_t box[5] = {64, (uint32_t)(ROW / 64), 1, 1, 1};
  if (encode_map(&ta, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  StemPoolParams p{};
  p.n = n;
  p.units = 4 * n;
  p.w_packed = w_packed;
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_pitch = out_pitch;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, StemPoolCfg::SMEM));
    attr_done = true;
  }
  int g = num_sms();
  if (g > p.units) g = p.units;
  launch_pdl_tpc(stem_pool_kernel, g, StemPoolCfg::THREADS, StemPoolCfg::SMEM, st, ta, p);
  return check_launch("stem_pool_kernel");
}

// ------------------------------------------------------------------ K1 fused into the stem (packed 224x224 uint8 crops)
static int stem_pool_u8_tc(const uint8_t* crops, const void* w_packed, const float* bias, int n, void* out, int64_t out_pitch,
                           cudaStream_t st) {
  AVCER_REQUIRE(out_pitch >= 64 && out_pitch % 8 == 0, "stem_pool_u8: out_pitch %lld must be a multiple of 8, at least 64", (long long)out_pitch);
  AVCER_REQUIRE(crops != nullptr && w_packed != nullptr && bias != nullptr && out != nullptr, "stem_pool_u8: null pointer");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(crops) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "stem_pool_u8: crops / w_packed / out must be 16-byte aligned");
  if (n == 0) return 0;
  StemPoolParams p{};
  p.n = n;
  p.units = 4 * n;
  p.w_packed = w_packed;
  p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_pitch = out_pitch;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(stem_pool_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, StemPoolU8Cfg::SMEM));
    attr_done = true;
  }
  int g = num_sms();
  if (g > p.units) g = p.units;
  launch_pdl_tpc(stem_pool_u8_kernel, g, StemPoolU8Cfg::THREADS, StemPoolU8Cfg::SMEM, st, crops, p);
  return check_launch("stem_pool_u8_kernel");
}

// ------------------------------------------------------------------ tcgen05 attention (bf16, head dim 64, T <= 208)
int attention_tc5(const void* qkv, int n, int t, int heads, float scale, void* out, cudaStream_t st) {
  using Cfg = Att5Cfg;
  AVCER_REQUIRE(t >= 1 && t <= Cfg::MAXT, "attention(tcgen05): T=%d exceeds %d", t, Cfg::MAXT);
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "attention(tcgen05): qkv / out must be 16-byte aligned");
  if (n == 0) return 0;
  Att5Params p{};
  p.n = n; p.t = t; p.heads = heads;
  p.npad = (t + 15) / 16 * 16;
  p.mtiles = (t + 127) / 128;
  p.units = n * heads;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  CUtensorMap tq, tkv;
  const uint64_t cols = (uint64_t)3 * heads * 64;
  uint64_t dims[3] = {cols, (uint64_t)t, (uint64_t)n};          // (column, token, window): tokens >= T are zero-filled
  uint64_t strides[2] = {cols * 2, (uint64_t)t * cols * 2};
  uint32_t boxq[3] = {64u, 128u, 1u};
  uint32_t boxkv[3] = {64u, (uint32_t)Cfg::MAXT, 1u};
  if (encode_map(&tq, qkv, 3, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (encode_map(&tkv, qkv, 3, dims, strides, boxkv, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  static bool attr_done = false;
  if (!attr_done) {
    AVCER_CUDA(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  int g = num_sms();
  if (g > p.units) g = p.units;
  launch_pdl_tpc(attention_tc5_kernel, g, Cfg::THREADS, Cfg::SMEM, st, tq, tkv, p);
  return check_launch("attention_tc5_kernel");
}

// ------------------------------------------------------------------ SIMT fp32 path
struct SimtParams {
  const float* a;
  long long a_dim[5], a_stride[5];
  const float* wt;
  const float* bias;
  const float* residual;
  float* out;
  long long out_stride[3], res_stride[3];
  int W, H, NB, cin, cout, taps_w, taps_h, off_w, off_h, tap_h_in_dim4, group_cin_shift, act, res_after_act;
  long long M;
  int Ktot;
  int vec4;   // cin % 4 == 0 and every address 16-byte aligned
};

__global__ void __launch_bounds__(256) simt_contract_kernel(const SimtParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int cshift = (n0 / 64) * p.group_cin_shift;

  // loader mapping: 64 rows x 4 quads of k
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;
  const long long m = m0 + lrow;
  const bool mvalid = m < p.M;
  int ow = 0, oh = 0, on = 0;
  if (mvalid) {
    ow = (int)(m % p.W);
    oh = (int)((m / p.W) % p.H);
    on = (int)(m / ((long long)p.W * p.H));
  }
  const int bcol = n0 + lrow;
  const bool bvalid = bcol < p.cout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;
  for (int k0 = 0; k0 < p.Ktot; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int kk = k0 + lk;
    if (mvalid && kk < p.Ktot) {
      if (p.vec4) {
        const int tap = kk / p.cin, c = kk % p.cin;
        const int tyy = tap / p.taps_w, txx = tap % p.taps_w;
        const long long cw = ow + p.off_w + txx;
        const long long ch = oh + p.off_h + (p.tap_h_in_dim4 ? 0 : tyy);
        const long long ct = p.tap_h_in_dim4 ? tyy : 0;
        const long long cc = c + cshift;
        if (cw >= 0 && cw < p.a_dim[1] && ch >= 0 && ch < p.a_dim[2] && on < p.a_dim[3] && ct < p.a_dim[4] &&
            cc + 3 < p.a_dim[0]) {
          const float4 v = *reinterpret_cast<const float4*>(p.a + cc + cw * p.a_stride[1] + ch * p.a_stride[2] +
                                                            on * p.a_stride[3] + ct * p.a_stride[4]);
          av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kk + j;
          if (k < p.Ktot) {
            const int tap = k / p.cin, c = k % p.cin;
            const int tyy = tap / p.taps_w, txx = tap % p.taps_w;
            const long long cw = ow + p.off_w + txx;
            const long long ch = oh + p.off_h + (p.tap_h_in_dim4 ? 0 : tyy);
            const long long ct = p.tap_h_in_dim4 ? tyy : 0;
            const long long cc = c + cshift;
            if (cw >= 0 && cw < p.a_dim[1] && ch >= 0 && ch < p.a_dim[2] && on < p.a_dim[3] && ct < p.a_dim[4] &&
                cc < p.a_dim[0])
              av[j] = p.a[cc + cw * p.a_stride[1] + ch * p.a_stride[2] + on * p.a_stride[3] + ct * p.a_stride[4]];
          }
        }
      }
    }
    if (bvalid && kk < p.Ktot) {
      if (p.vec4) {
        const float4 v = *reinterpret_cast<const float4*>(p.wt + (long long)bcol * p.Ktot + kk);
        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (kk + j < p.Ktot) bv[j] = p.wt[(long long)bcol * p.Ktot + kk + j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = av[j];
      Bs[lk + j][lrow] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long mm = m0 + ty * 4 + i;
    if (mm >= p.M) continue;
    const int w = (int)(mm % p.W);
    const int h = (int)((mm / p.W) % p.H);
    const int n = (int)(mm / ((long long)p.W * p.H));
    const long long oo = w * p.out_stride[0] + h * p.out_stride[1] + n * p.out_stride[2];
    const long long ro = w * p.res_stride[0] + h * p.res_stride[1] + n * p.res_stride[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[co];
      if (p.residual && !p.res_after_act) v += p.residual[ro + co];
      if (p.act == ACT_RELU) v = v < 0.f ? 0.f : v;
      else if (p.act == ACT_GELU) v = gelu_erf(v);
      if (p.residual && p.res_after_act) v += p.residual[ro + co];
      p.out[oo + co] = v;
    }
  }
}

static int contract_simt(const avcer_contract_desc* d, cudaStream_t st) {
  AVCER_REQUIRE(d->a_step <= 1, "contract(fp32): a_step is a bf16-path feature");
  SimtParams p{};
  p.a = static_cast<const float*>(d->a);
  for (int i = 0; i < 5; ++i) { p.a_dim[i] = d->a_dim[i]; p.a_stride[i] = d->a_stride[i]; }
  p.wt = static_cast<const float*>(d->wt);
  p.bias = d->bias;
  p.residual = static_cast<const float*>(d->residual);
  p.out = static_cast<float*>(d->out);
  for (int i = 0; i < 3; ++i) { p.out_stride[i] = d->out_stride[i]; p.res_stride[i] = d->res_stride[i]; }
  p.W = d->W; p.H = d->H; p.NB = d->NB; p.cin = d->cin; p.cout = d->cout;
  p.taps_w = d->taps_w; p.taps_h = d->taps_h; p.off_w = d->off_w; p.off_h = d->off_h;
  p.tap_h_in_dim4 = d->tap_h_in_dim4; p.group_cin_shift = d->group_cin_shift; p.act = d->act; p.res_after_act = d->res_after_act;
  p.M = (long long)d->W * d->H * d->NB;
  p.Ktot = d->taps_w * d->taps_h * d->cin;
  AVCER_REQUIRE(d->group_cin_shift == 0 || d->cout % 64 == 0, "contract(f32): grouped conv needs cout %% 64 == 0");
  bool v = (d->cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(d->a) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(d->wt) & 15) == 0) && (d->group_cin_shift % 4 == 0);
  for (int i = 1; i < 5; ++i) v = v && (d->a_stride[i] % 4 == 0);
  p.vec4 = v ? 1 : 0;
  if (p.M == 0) return 0;
  const long long gx = (p.M + 63) / 64;
  AVCER_REQUIRE(gx < (1ll << 31), "contract(f32): M too large");
  dim3 grid((unsigned)gx, (unsigned)((d->cout + 63) / 64));
  g_last_kernel = KID_SIMT;
  simt_contract_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("simt_contract_kernel");
}

}  // namespace avcer

extern "C" int avcer_contract(const avcer_contract_desc* d, void* stream) {
  using namespace avcer;
  AVCER_REQUIRE(d != nullptr, "contract: null descriptor");
  AVCER_REQUIRE(d->W > 0 && d->H > 0 && d->NB >= 0 && d->cin > 0 && d->cout > 0 && d->taps_w > 0 && d->taps_h > 0,
                "contract: bad geometry");
  if (d->dtype == AVCER_BF16) return contract_tc(d, as_stream(stream));
  if (d->dtype == AVCER_F32) return contract_simt(d, as_stream(stream));
  return set_error("contract: unknown dtype %d", d->dtype);
}

extern "C" int avcer_stem_pool(const void* x_padded, const void* w_packed, const float* bias, int n, void* out, void* stream) {
  using namespace avcer;
  AVCER_REQUIRE(n >= 0, "stem_pool: negative batch");
  return stem_pool_tc(x_padded, w_packed, bias, n, out, 64, as_stream(stream));
}

extern "C" int avcer_stem_pool_ld(const void* x_padded, const void* w_packed, const float* bias, int n, void* out, int64_t out_pitch,
                                  void* stream) {
  using namespace avcer;
  AVCER_REQUIRE(n >= 0, "stem_pool: negative batch");
  return stem_pool_tc(x_padded, w_packed, bias, n, out, out_pitch, as_stream(stream));
}

extern "C" int avcer_stem_pool_u8(const uint8_t* crops, const void* w_packed, const float* bias, int n, void* out, int64_t out_pitch,
                                  void* stream) {
  using namespace avcer;
  AVCER_REQUIRE(n >= 0, "stem_pool_u8: negative batch");
  return stem_pool_u8_tc(crops, w_packed, bias, n, out, out_pitch, as_stream(stream));
}

// ------------------------------------------------------------------ wav2vec2 conv0 + LayerNorm + GELU on the tensor cores
extern "C" int avcer_w2v_conv0_tc(const float* x, int n, int t_in, const void* w_packed, const float* ln_g, const float* ln_b,
                                  float eps, void* y, int64_t y_pitch_rows, void* stream) {
  using namespace avcer;
  AVCER_REQUIRE(t_in >= 10 && n >= 0, "w2v_conv0_tc: bad shape n=%d t_in=%d", n, t_in);
  const int t_out = (t_in - 10) / 5 + 1;
  AVCER_REQUIRE(y_pitch_rows >= t_out, "w2v_conv0_tc: y pitch too small");
  AVCER_REQUIRE(x != nullptr && w_packed != nullptr && ln_g != nullptr && ln_b != nullptr && y != nullptr, "w2v_conv0_tc: null pointer");
  AVCER_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "w2v_conv0_tc: w_packed / y must be 16-byte aligned");
  if (n == 0) return 0;
  CUtensorMap ty;
  uint64_t dims[5] = {512, (uint64_t)t_out, (uint64_t)n, 1, 1};
  uint64_t strides[4] = {512 * 2, (uint64_t)y_pitch_rows * 512 * 2, (uint64_t)1 << 30, (uint64_t)1 << 30};
  uint32_t box[5] = {64u, 32u, 1u, 1u, 1u};
  if (encode_map(&ty, y, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  Conv0Params p{};
  p.x = x; p.w_packed = w_packed; p.gamma = ln_g; p.beta = ln_b;
  p.n = n; p.t_in = t_in; p.t_out = t_out;
  p.tiles_per_row = (t_out + 127) / 128;
  p.num_tiles = n * p.tiles_per_row;
  p.eps = eps;
  // 8 epilogue warps (2 per TMEM lane quarter) by default.  AVCER_CONV0_G=4 selects 16 (96 registers per thread): measured
  // 352 vs 336 us per 64 windows -- more warps do not help, both the MUFU and the FMA pipe sit at ~50 % (ncu) whatever the
  // warp count, and prefetching the TMEM loads gained 5 %.
  static const int groups = getenv("AVCER_CONV0_G") ? atoi(getenv("AVCER_CONV0_G")) : 2;
  int g = num_sms();
  if (g > p.num_tiles) g = p.num_tiles;
  static bool attr_done[2] = {false, false};
  if (groups == 2) {
    if (!attr_done[0]) {
      AVCER_CUDA(cudaFuncSetAttribute(w2v_conv0_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv0Cfg<2>::SMEM));
      attr_done[0] = true;
    }
    launch_pdl_tpc(w2v_conv0_tc_kernel<2>, g, Conv0Cfg<2>::THREADS, Conv0Cfg<2>::SMEM, as_stream(stream), ty, p);
  } else {
    if (!attr_done[1]) {
      AVCER_CUDA(cudaFuncSetAttribute(w2v_conv0_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv0Cfg<4>::SMEM));
      attr_done[1] = true;
    }
    launch_pdl_tpc(w2v_conv0_tc_kernel<4>, g, Conv0Cfg<4>::THREADS, Conv0Cfg<4>::SMEM, as_stream(stream), ty, p);
  }
  return check_launch("w2v_conv0_tc_kernel");
}

extern "C" int avcer_last_contract_kernel(void) { return avcer::g_last_kernel; }
