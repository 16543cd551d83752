// bf16 GEMM on the 5th-generation tensor cores: C = act(A W^T + bias), fp32 accumulation in TMEM.
//
//   A (M,K) bf16, pitch lda  — activations (rows = tokens)       -> UMMA operand A, K-major
//   W (N,K) bf16, pitch ldw  — nn.Linear weight as stored (out,in) -> UMMA operand B, K-major
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0        TMA producer: cp.async.bulk.tensor 128x64 (A) and 256x64 (W) bf16 boxes, 128-byte
//                 swizzle, into a 4-stage shared-memory ring guarded by full/empty mbarriers
//   warp 1        MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256,
//                 K=16) x4 per stage; tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..9    epilogue: tcgen05.ld the 128x256 fp32 accumulator (two warps per TMEM lane quarter),
//                 bias + activation in registers, then each warp stages its 32-row x 128-byte chunk in
//                 128B-swizzled shared memory and ONE lane writes it with a TMA bulk tensor store
//                 (full-line writes, M/N tails clipped by the tensor map).  Round-1 profile: per-row
//                 16-byte st.global from registers cost 32 sectors/request and made K=512 GEMMs
//                 epilogue-bound at 17% of peak (profiles/r01_gemm_epilogue.md).
// The accumulator is double-buffered in TMEM (2 x 256 columns = all 512) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Out-of-range rows/columns/K are zero-filled by TMA on load and
// masked on store, so M, N, K need no padding beyond the 16-byte pitch alignment TMA requires.
//
// Replaces every nn.Linear on the denoiser path in bf16 mode (see include/tcdiff_b200.h: tcd_gemm).
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
// variants made the linear1 epilogue 2.6x slower than the MMAs (profiles/r01_gemm_epilogue.md).
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

template <typename OutT, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tc_kernel(
    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
    const __grid_constant__ CUtensorMap tmap_c, int use_tma_store, const float* __restrict__ bias, int act,
    OutT* __restrict__ C, int64_t ldc, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;   // 1024-byte aligned staging slots
  const uint32_t bar_base = epi_base + EPI_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + EPI_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    if (use_tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate all 512 TMEM columns, address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m0);
          tma_load_2d(sa + A_STAGE_BYTES, &tmap_b, full_bar(stage), kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);       // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr>>4) units
            tc_mma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (kb | k) != 0);
          }
          tc_commit(empty_bar(stage));             // smem stage is free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));                  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                   // TMEM lanes [32*quarter, +32) are visible to this warp
    const int half = ew >> 2;                       // which 128-column half of the accumulator
    constexpr int CH = 128 / (int)sizeof(OutT);     // columns per 128-byte staging row: 32 fp32 / 64 bf16
    const uint32_t slot = epi_base + (uint32_t)(ew * EPI_SLOT_BYTES);
    const bool vec_ok = (ldc % (16 / (int)sizeof(OutT)) == 0) && ((uintptr_t)C % 16 == 0);
    const bool bias_vec = ((uintptr_t)bias % 16) == 0;          // tile column offsets are multiples of 32
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row0 = m0 + quarter * 32;
      const int row = row0 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += CH) {
        const int col0 = n0 + half * (BN / 2) + c;
        if (col0 >= N) break;                       // warp-uniform
        float v[CH];
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          uint32_t raw[32];
          tc_ld32(taddr + (uint32_t)(c + 32 * q), raw);
          const int cq = col0 + 32 * q;
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
#pragma unroll
          for (int j = 0; j < 32; ++j) v[32 * q + j] = epi_act<ACT>(__uint_as_float(raw[j]) + bv[j], act);
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
              tma_store_2d(&tmap_c, slot, col0, row0);
              tma_store_commit();
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    if (use_tma_store && lane == 0) tma_store_wait_all();   // smem must outlive the last bulk store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major (rows, cols) tensor with pitch ld elements; box = (box_rows, 128 bytes of columns), 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return TCD_ERR_CUDA; }
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r, (long long)rows, (long long)cols, (long long)ld); return TCD_ERR_CUDA; }
  return TCD_OK;
}

// 3-D bf16 tensor (col, row, batch): pitch ld, batch stride in elements; box = (64 cols, box_rows, 1), 128B swizzle
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld,
                      int64_t batch_stride, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return TCD_ERR_CUDA; }
  if (batches <= 1 && batch_stride < ld) batch_stride = ld * rows;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batches < 1 ? 1 : batches)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d (cols=%lld rows=%lld batches=%lld ld=%lld bs=%lld)", (int)r, (long long)cols, (long long)rows, (long long)batches, (long long)ld, (long long)batch_stride); return TCD_ERR_CUDA; }
  return TCD_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename OutT, int ACT>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                     const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<OutT, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) { set_error("gemm_bf16_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tc_kernel<OutT, ACT><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ta, tb, tc, use_tma_store, bias, act, (OutT*)C, ldc, M, N, K);
  return check_launch("gemm_bf16_tc");
}

int gemm_bf16_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act, int out_dtype,
                 void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  TCD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && lda % 8 == 0 && ldw % 8 == 0,
              "tcd_gemm(bf16): A/W base and pitch must be 16-byte aligned (lda=%lld ldw=%lld)", (long long)lda, (long long)ldw);
  TCD_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "tcd_gemm(bf16): dimension too large");
  CUtensorMap ta, tb, tc;
  int rc = make_tmap_2d(&ta, A, M, K, lda, BM, false);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, N, K, ldw, BN, false);
  if (rc) return rc;
  // the bulk-store epilogue needs a 16-byte aligned output base and pitch; otherwise (final_layer, N = ldc = 151)
  // the kernel falls back to per-row stores from registers
  const bool f32 = out_dtype == TCD_F32;
  const int es = f32 ? 4 : 2;
  const int use_tma_store = ((uintptr_t)C % 16 == 0) && ((ldc * es) % 16 == 0);
  if (use_tma_store) {
    rc = make_tmap_2d(&tc, C, M, N, ldc, 32, f32);
    if (rc) return rc;
  } else {
    tc = ta;
  }
#define TCD_LAUNCH(OUT, ACTV) launch_tc<OUT, ACTV>(ta, tb, tc, use_tma_store, bias, act, C, ldc, (int)M, (int)N, (int)K, st)
  if (f32) {
    switch (act) {
      case TCD_ACT_NONE: return TCD_LAUNCH(float, TCD_ACT_NONE);
      case TCD_ACT_RELU: return TCD_LAUNCH(float, TCD_ACT_RELU);
      case TCD_ACT_GELU: return TCD_LAUNCH(float, TCD_ACT_GELU);
      default: return TCD_LAUNCH(float, ACT_RUNTIME);
    }
  }
  switch (act) {
    case TCD_ACT_NONE: return TCD_LAUNCH(__nv_bfloat16, TCD_ACT_NONE);
    case TCD_ACT_RELU: return TCD_LAUNCH(__nv_bfloat16, TCD_ACT_RELU);
    case TCD_ACT_GELU: return TCD_LAUNCH(__nv_bfloat16, TCD_ACT_GELU);
    default: return TCD_LAUNCH(__nv_bfloat16, ACT_RUNTIME);
  }
#undef TCD_LAUNCH
}

}  // namespace tcd
