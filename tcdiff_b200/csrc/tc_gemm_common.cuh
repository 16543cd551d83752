// Shared tile constants and inline-PTX wrappers of the tcgen05 GEMM kernels (gemm_tc.cu: cta_group::1,
// gemm_tc2.cu: cta_group::2).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tcd {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BN = 256;       // UMMA N
constexpr int BK = 64;        // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = (2 + EPI_WARPS) * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int TMEM_COLS = 512;               // two 256-column fp32 accumulators
constexpr int EPI_SLOT_BYTES = 32 * 128;      // per epilogue warp: 32 rows x 128 bytes (32 fp32 or 64 bf16 columns)
constexpr int EPI_BYTES = EPI_WARPS * EPI_SLOT_BYTES;
constexpr size_t GEMM_SMEM = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/;

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();  // a lost arrive becomes an error instead of a hung GPU
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- converged single-issuer variants ------------------------------------------------------------------------
// A producer / MMA loop written as `if (lane == 0) { loop }` is divergent code for ptxas: descriptors and barrier
// addresses live in vector registers and every UTCHMMA / UTCBAR / UTMALDG gets an R2UR + ELECT "waterfall" loop
// (~25 dependent instructions per tcgen05.mma, measured r01: the attention kernel was paced by its MMA thread).
// The `_p` wrappers let ALL lanes run the warp-uniform loop (operands stay in uniform registers) and apply the
// elected-lane predicate inside the asm statement, so the C++ control flow never diverges.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_expect_tx_p(uint32_t leader, uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes), "r"(leader) : "memory");
}
__device__ __forceinline__ void tma_load_2d_p(uint32_t leader, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_commit_p(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_p(uint32_t leader, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024 B (8 rows)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// GELU (exact-erf form, F.gelu default) for the bf16 epilogue.  gelu(v) = v * Phi(v) with
// Phi(v) = 0.5 * erfc(-v / sqrt(2)); for z = |v|/sqrt(2), Abramowitz-Stegun 7.1.26 gives
// erfc(z) = poly(t) * exp(-z^2), t = 1/(1 + p z), |err| < 1.5e-7 (far below the bf16 output rounding of 2^-9).
// Two MUFU ops (rcp.approx, ex2.approx) + ~10 FMA-pipe ops per element; the IEEE-rounded __frcp_rn / erff()
// variants made the linear1 epilogue 2.6x slower than the MMAs (profiles/r01_gemm_epilogue.md).  A rational erf with ONE
// MUFU op and 13 packed FMA-pipe instructions per pair measured SLOWER (r02: 112.6 vs 104.5 us for 96000 x 1024 x 512,
// profiles/r02_ab.md): the epilogue is bound by issue slots, not by the MUFU.
__device__ __forceinline__ float gelu_fast(float v) {
  const float z = fabsf(v) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float half_erfc = 0.5f * p * e;                    // 0.5 * erfc(|v|/sqrt2) = Phi(-|v|)
  const float phi = v >= 0.f ? 1.0f - half_erfc : half_erfc;
  return v * phi;
}
// Two elements at a time on the packed fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2): the same operations in the same order
// and rounding as gelu_fast, hence bit-identical results, with half the FMA-pipe instructions — the linear1 epilogue
// is FMA/MUFU-bound (~14 FMA-pipe + 2 MUFU instructions per element against 4096 MMA cycles per 128x256x512 tile).
__device__ __forceinline__ float2 gelu_fast2(float2 v) {
  const float2 z = __fmul2_rn(make_float2(fabsf(v.x), fabsf(v.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  p = __fmul2_rn(p, t);
  const float2 arg = __fmul2_rn(__fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f)), z);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
  const float2 h = __fmul2_rn(__fmul2_rn(p, make_float2(0.5f, 0.5f)), e);
  const float2 phi = make_float2(v.x >= 0.f ? 1.0f - h.x : h.x, v.y >= 0.f ? 1.0f - h.y : h.y);
  return __fmul2_rn(v, phi);
}

// ACT is a compile-time constant for the hot instantiations (none / relu / gelu) so the 32-element epilogue
// loop is straight-line code; ACT_RUNTIME serves the tiny Mish / SiLU GEMMs of the conditioning path.
// (Round-1 profile: a runtime switch inlined per element produced ~5000 SASS instructions per chunk.)
constexpr int ACT_RUNTIME = -1;
template <int ACT>
__device__ __forceinline__ float epi_act(float v, int act) {
  if constexpr (ACT == TCD_ACT_NONE) return v;
  else if constexpr (ACT == TCD_ACT_RELU) return fmaxf(v, 0.f);
  else if constexpr (ACT == TCD_ACT_GELU) return gelu_fast(v);
  else {
    switch (act) {
      case TCD_ACT_RELU: return fmaxf(v, 0.f);
      case TCD_ACT_GELU: return gelu_fast(v);
      case TCD_ACT_MISH: return act_mish(v);
      case TCD_ACT_SILU: return act_silu(v);
      case TCD_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
      default: return v;
    }
  }
}

template <typename OutT>
__device__ __forceinline__ void store_chunk(OutT* dst, const float (&v)[32], int ncols, bool vec_ok);

template <>
__device__ __forceinline__ void store_chunk<float>(float* dst, const float (&v)[32], int ncols, bool vec_ok) {
  if (vec_ok && ncols == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = v[j];
  }
}
template <>
__device__ __forceinline__ void store_chunk<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[32], int ncols,
                                                           bool vec_ok) {
  if (vec_ok && ncols == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v[j], v[j + 1]), b = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[j + 4], v[j + 5]), d = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
      u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
      *reinterpret_cast<uint4*>(dst + j) = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
  }
}


// Drain this warp's 32-row x 128-column part of a 128x256 fp32 accumulator: tcgen05.ld, bias + activation in
// registers, then either (use_tma_store) stage 32 x 128-byte chunks in 128B-swizzled smem and write them with one
// TMA bulk store each, or fall back to per-row stores when the output pitch is not 16-byte aligned.
//   taddr: TMEM address of (first lane of the warp's quarter, first of its 128 columns)
//   row0:  global row of lane 0;  colbase: global column of the first of the 128 columns
template <typename OutT, int ACT, bool ROWSTORE = false>
__device__ __forceinline__ void epilogue_drain(uint32_t taddr, int row0, int colbase, int lane, uint32_t slot,
                                               const CUtensorMap* tmap_c_ptr, int use_tma_store,
                                               const float* __restrict__ bias, bool bias_vec, int act,
                                               OutT* __restrict__ C, int64_t ldc, bool vec_ok, int M, int N) {
  constexpr int CH = 128 / (int)sizeof(OutT);     // columns per 128-byte staging row: 32 fp32 / 64 bf16
  const int row = row0 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += CH) {
        const int col0 = colbase + c;
        if (col0 >= N) break;                       // warp-uniform
        float v[CH];
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          uint32_t raw[32];
          tc_ld32(taddr + (uint32_t)(c + 32 * q), raw);
          const int cq = col0 + 32 * q;
          // bias == NULL (most decoder GEMMs) is its own straight-line case.  ncu r02 (96 000 x 512 x 512, bias = NULL): the
          // single predicated path issued 141 ISETP + 131 R2UR + 65 VIADD + 63 predicated-off LDG + 64 FADD per chunk even
          // without a bias — 589 instructions per 64-column chunk.  (The GELU instantiation keeps the one path it always
          // takes, aligned bias: splitting it as well made the linear1 GEMM 8 % slower, 104.5 -> 112.7 us.)
          bool no_bias = false;
          if constexpr (ACT != TCD_ACT_GELU) no_bias = bias == nullptr;         // warp-uniform
          if (no_bias) {
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[32 * q + j] = epi_act<ACT>(__uint_as_float(raw[j]), act);
          } else {
            float bv[32];
            if (bias != nullptr && cq + 32 <= N && bias_vec) {        // warp-uniform; 8 broadcast 16-byte loads
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 t = __ldg(reinterpret_cast<const float4*>(bias + cq + j));
                bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) bv[j] = (bias != nullptr && cq + j < N) ? __ldg(bias + cq + j) : 0.f;
            }
            tc_wait_ld();
            if constexpr (ACT == TCD_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float2 r = gelu_fast2(__fadd2_rn(make_float2(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1])),
                                                       make_float2(bv[j], bv[j + 1])));
                v[32 * q + j] = r.x;
                v[32 * q + j + 1] = r.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[32 * q + j] = epi_act<ACT>(__uint_as_float(raw[j]) + bv[j], act);
            }
          }
        }
        if (use_tma_store) {
          if (row0 < M) {                           // warp-uniform
            if (lane == 0) tma_store_wait_read();   // previous bulk store has finished reading this slot
            __syncwarp();
            // row `lane` of the 32 x 128-byte box, 16-byte chunk j stored at j ^ (row & 7) (SWIZZLE_128B)
            const uint32_t rbase = slot + (uint32_t)(lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t w0, w1, w2, w3;
              if constexpr (sizeof(OutT) == 4) {
                w0 = __float_as_uint(v[4 * j]); w1 = __float_as_uint(v[4 * j + 1]);
                w2 = __float_as_uint(v[4 * j + 2]); w3 = __float_as_uint(v[4 * j + 3]);
              } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                w0 = *reinterpret_cast<uint32_t*>(&p0); w1 = *reinterpret_cast<uint32_t*>(&p1);
                w2 = *reinterpret_cast<uint32_t*>(&p2); w3 = *reinterpret_cast<uint32_t*>(&p3);
              }
              sts128(rbase + (uint32_t)(((j ^ lane) & 7) << 4), w0, w1, w2, w3);
            }
            fence_proxy_async();                    // generic-proxy smem writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tmap_c_ptr, slot, col0, row0);
              tma_store_commit();
            }
          }
        } else if (ROWSTORE) {
          // Unaligned output pitch (final_layer: N = ldc = 151), rows transposed through the warp's staging slot: every
          // store instruction writes 32 consecutive columns of ONE row (coalesced) instead of one column of 32 rows
          // (32 sectors per request: 123 us for the 96 000 x 151 x 512 head GEMM of the c2 sampler, r01).
          // Word (row i, column j) of the 32 x 32 chunk lives at i*32 + (j ^ i): conflict-free both ways.
          // ROWSTORE is a compile-time constant: the other instantiations keep their machine code.
#pragma unroll
          for (int q = 0; q < CH / 32; ++q) {
            const int cq = col0 + 32 * q;
            if (cq < N && row0 < M) {               // warp-uniform
              const int ncols = min(32, N - cq), nrows = min(32, M - row0);
              __syncwarp();                         // the previous read-back of the slot is complete
#pragma unroll
              for (int j = 0; j < 32; ++j)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + 4u * (uint32_t)(lane * 32 + (j ^ lane))), "f"(v[32 * q + j]) : "memory");
              __syncwarp();
              for (int i = 0; i < nrows; ++i) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(slot + 4u * (uint32_t)(i * 32 + (lane ^ i))) : "memory");
                if (lane < ncols) C[(int64_t)(row0 + i) * ldc + cq + lane] = Conv<OutT>::to(t);
              }
            }
          }
        } else if (row < M) {
#pragma unroll
          for (int q = 0; q < CH / 32; ++q) {
            const int cq = col0 + 32 * q;
            if (cq < N) {
              float t[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) t[j] = v[32 * q + j];
              store_chunk<OutT>(C + (int64_t)row * ldc + cq, t, min(32, N - cq), vec_ok);
            }
          }
        }
      }
}

}  // namespace tcd
