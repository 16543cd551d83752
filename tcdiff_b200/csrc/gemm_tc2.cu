// bf16 GEMM, 2-CTA variant: C = act(A W^T + bias) with tcgen05.mma.cta_group::2 (UMMA M = 256).
//
// A CTA pair (cluster of 2, same TPC) owns one 256 x 256 output tile per scheduling step.  Each CTA loads its own
// 128 rows of A and only HALF (128 of 256 rows) of the W tile; the pair's tensor cores read both halves, so the
// L2->SM operand traffic per FLOP drops from (128+256)x64 to (128+128)x64 elements per 128x256x64 MACs
// (85 -> 128 FLOP per byte) — the 1-CTA kernel (gemm_tc.cu) sits at the ~12 TB/s L2->SM ceiling on the decoder's
// K = 512 shapes.  Per CTA: 6-stage x 32 KiB TMA ring, 2 x 256-column TMEM accumulators (own 128 rows of D),
// the same 8-warp TMA-store epilogue as the 1-CTA kernel.
//
// Protocol (CUTLASS 2-SM scheme restated):
//   * both CTAs issue `cp.async.bulk.tensor...cta_group::2` loads into their own smem; the transaction bytes of BOTH
//     land on the LEADER's (cluster rank 0) full barrier (barrier address with the peer bit cleared), which expects
//     2 x 32 KiB per stage;
//   * only the leader issues tcgen05.mma.cta_group::2; `tcgen05.commit...multicast::cluster` (mask 0b11) releases the
//     smem stage in BOTH CTAs and publishes the accumulator to BOTH CTAs' epilogue warps;
//   * the epilogue warps of both CTAs arrive on the leader's tmem-empty barrier (remote arrive through mapa);
//   * TMEM is allocated/freed with cta_group::2 by warp 1 of both CTAs; cluster barriers bracket setup and teardown.
#include "tc_gemm_common.cuh"

namespace tcd {

#ifdef TCD_GEMM_DEBUG
// Development aid (A/B builds only, -DTCD_GEMM_DEBUG): cycle counters of the CTA-pair GEMM's roles, summed over all CTAs.
//   [0] MMA warp: waiting for an accumulator (tempty)   [1] MMA warp: waiting for operands (full)   [2] MMA warp: whole loop
//   [3] epilogue warp 2: waiting for the MMAs (tfull)   [4] epilogue warp 2: draining               [5] tiles (leader CTAs)
__device__ unsigned long long g_gemm_dbg[8];
#define TCD_DBG_CLK() clock64()
#define TCD_DBG_ADD(i, v) atomicAdd(&g_gemm_dbg[i], (unsigned long long)(v))
#else
#define TCD_DBG_CLK() 0ll
#define TCD_DBG_ADD(i, v)
#endif

int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32);
int num_sms();

constexpr int STAGES2 = 6;
constexpr int A2_BYTES = 128 * BK * 2;            // 16 KiB: this CTA's 128 rows of A
constexpr int B2_BYTES = 128 * BK * 2;            // 16 KiB: this CTA's half of the 256-row W tile
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr size_t GEMM2_SMEM = 1024 + (size_t)STAGES2 * STAGE2_BYTES + EPI_BYTES + 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;       // clears the CTA-pair rank bit of a shared::cluster address
// kind::f16, D f32, A/B bf16 K-major, N = 256, M = 256 (cta_group::2)
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar) {   // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_p(uint32_t leader, uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                  int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_commit_mc2_p(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"((uint16_t)3), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16_p(uint32_t leader, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t target_cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar), "r"(target_cta) : "memory");
}

// CONV: converged producer / MMA issue loops (tc_gemm_common.cuh, `_p` wrappers); 0 keeps the lane-0 loops.
// ROWSTORE: unaligned output pitch -> rows transposed through shared memory (epilogue_drain); float / no activation only.
template <typename OutT, int ACT, int CONV, bool ROWSTORE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tc2_kernel(
    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
    const __grid_constant__ CUtensorMap tmap_c, int use_tma_store, const float* __restrict__ bias, int act,
    OutT* __restrict__ C, int64_t ldc, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES2 * STAGE2_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES2 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES2 + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES2 * STAGE2_BYTES + EPI_BYTES + 8 * (2 * STAGES2 + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0 = leader of the pair
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_m = (M + 255) / 256, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    if (use_tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // same warp id in both CTAs, same smem destination
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();        // barrier inits + TMEM allocation visible to both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_sync();                // everything above overlaps the previous kernel's tail; A, bias and C only from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if constexpr (CONV != 0) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m0 = (tile / tiles_n) * 256 + (int)rank * 128;
        const int nb = (tile % tiles_n) * BN + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_expect_tx_p(leader, full_bar(stage), 2 * STAGE2_BYTES);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint32_t lbar = full_bar(stage) & kPeerMask;
          tma_load_2d_2sm_p(leader, sa, &tmap_a, lbar, kb * BK, m0);
          tma_load_2d_2sm_p(leader, sa + A2_BYTES, &tmap_b, lbar, kb * BK, nb);
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m0 = (tile / tiles_n) * 256 + (int)rank * 128;
        const int nb = (tile % tiles_n) * BN + (int)rank * 128;      // this CTA's half of the W tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);                  // own barrier, released by the multicast commit
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE2_BYTES);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint32_t lbar = full_bar(stage) & kPeerMask;        // the leader's barrier
          tma_load_2d_2sm(sa, &tmap_a, lbar, kb * BK, m0);
          tma_load_2d_2sm(sa + A2_BYTES, &tmap_b, lbar, kb * BK, nb);
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (CONV != 0 && rank == 0) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      long long d_te = 0, d_fu = 0;
      const long long d_t0 = TCD_DBG_CLK();
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        const long long c0 = TCD_DBG_CLK();
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        d_te += TCD_DBG_CLK() - c0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          const long long c1 = TCD_DBG_CLK();
          mbar_wait(full_bar(stage), phase);
          d_fu += TCD_DBG_CLK() - c1;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + A2_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc_mma2_f16_p(leader, tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc2, (uint32_t)(kb | k));
          tc_commit_mc2_p(leader, empty_bar(stage));
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
        tc_commit_mc2_p(leader, tfull_bar(as));
      }
      if (lane == 0) { TCD_DBG_ADD(0, d_te); TCD_DBG_ADD(1, d_fu); TCD_DBG_ADD(2, TCD_DBG_CLK() - d_t0); TCD_DBG_ADD(5, it); }
    } else if (CONV == 0 && rank == 0 && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);                     // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);                        // both CTAs' TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + A2_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc_mma2_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc2, (kb | k) != 0);
          tc_commit_mc2(empty_bar(stage));
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
        tc_commit_mc2(tfull_bar(as));
      }
    }
  } else {
    // ===================== epilogue (8 warps, both CTAs: own 128 rows of D) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const uint32_t slot = epi_base + (uint32_t)(ew * EPI_SLOT_BYTES);
    const bool vec_ok = (ldc % (16 / (int)sizeof(OutT)) == 0) && ((uintptr_t)C % 16 == 0);
    const bool bias_vec = ((uintptr_t)bias % 16) == 0;
    int it = 0;
    long long d_tf = 0, d_dr = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int m0 = (tile / tiles_n) * 256 + (int)rank * 128, n0 = (tile % tiles_n) * BN;
      const long long e0 = TCD_DBG_CLK();
      mbar_wait(tfull_bar(as), aphase);
      const long long e1 = TCD_DBG_CLK();
      d_tf += e1 - e0;
      tc_fence_after();
      const int row0 = m0 + quarter * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
      epilogue_drain<OutT, ACT, ROWSTORE>(taddr, row0, n0 + half * (BN / 2), lane, slot, &tmap_c, use_tma_store, bias, bias_vec, act,
                                C, ldc, vec_ok, M, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(as));
        else mbar_arrive_remote(tempty_bar(as), 0);
      }
      d_dr += TCD_DBG_CLK() - e1;
    }
    if (ew == 0 && lane == 0 && rank == 0) { TCD_DBG_ADD(3, d_tf); TCD_DBG_ADD(4, d_dr); }
    if (use_tma_store && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();        // no CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Unaligned output pitch (no TMA store, final_layer: N = ldc = 151): rows are transposed through shared memory for
// coalesced stores in the float / no-activation instantiation (r01: head GEMM 123 -> 69 us, same bits).

template <typename OutT, int ACT, int CONV, bool ROWSTORE>
static int launch_tc2r(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                       const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc2_kernel<OutT, ACT, CONV, ROWSTORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM2_SMEM);
    if (e != cudaSuccess) { set_error("gemm_bf16_tc2: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int tiles = ((M + 255) / 256) * ((N + BN - 1) / BN);
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaError_t le = launch_pdl(gemm_bf16_tc2_kernel<OutT, ACT, CONV, ROWSTORE>, dim3(2 * pairs), dim3(GEMM_THREADS), GEMM2_SMEM, st, tiles <= 2 * max_pairs, ta, tb, tc,
                              use_tma_store, bias, act, (OutT*)C, ldc, M, N, K);
  if (le != cudaSuccess) { set_error("gemm_bf16_tc2: launch: %s", cudaGetErrorString(le)); return TCD_ERR_CUDA; }
  return check_launch("gemm_bf16_tc2");
}

template <typename OutT, int ACT, int CONV>
static int launch_tc2v(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                      const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  if constexpr (sizeof(OutT) == 4 && ACT == TCD_ACT_NONE) {
    if (!use_tma_store)
      return launch_tc2r<OutT, ACT, CONV, true>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);
  }
  return launch_tc2r<OutT, ACT, CONV, false>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);
}

// Issue loops: converged in the CTA-pair kernel (r01 A/B: +2..10 %), lane-0 in the 1-CTA kernel (converged: -2 %).
template <typename OutT, int ACT>
static int launch_tc2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                      const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  return launch_tc2v<OutT, ACT, 1>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);
}

int gemm_bf16_tc2(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act, int out_dtype,
                  void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  CUtensorMap ta, tb, tc;
  int rc = make_tmap_2d(&ta, A, M, K, lda, 128, false);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, N, K, ldw, 128, false);        // half of the 256-row W tile per CTA
  if (rc) return rc;
  const bool f32 = out_dtype == TCD_F32;
  const int es = f32 ? 4 : 2;
  const int use_tma_store = ((uintptr_t)C % 16 == 0) && ((ldc * es) % 16 == 0);
  if (use_tma_store) {
    rc = make_tmap_2d(&tc, C, M, N, ldc, 32, f32);
    if (rc) return rc;
  } else {
    tc = ta;
  }
#define TCD_LAUNCH2(OUT, ACTV) launch_tc2<OUT, ACTV>(ta, tb, tc, use_tma_store, bias, act, C, ldc, (int)M, (int)N, (int)K, st)
  if (f32) {
    switch (act) {
      case TCD_ACT_NONE: return TCD_LAUNCH2(float, TCD_ACT_NONE);
      case TCD_ACT_RELU: return TCD_LAUNCH2(float, TCD_ACT_RELU);
      case TCD_ACT_GELU: return TCD_LAUNCH2(float, TCD_ACT_GELU);
      default: return TCD_LAUNCH2(float, ACT_RUNTIME);
    }
  }
  switch (act) {
    case TCD_ACT_NONE: return TCD_LAUNCH2(__nv_bfloat16, TCD_ACT_NONE);
    case TCD_ACT_RELU: return TCD_LAUNCH2(__nv_bfloat16, TCD_ACT_RELU);
    case TCD_ACT_GELU: return TCD_LAUNCH2(__nv_bfloat16, TCD_ACT_GELU);
    default: return TCD_LAUNCH2(__nv_bfloat16, ACT_RUNTIME);
  }
#undef TCD_LAUNCH2
}

}  // namespace tcd

#ifdef TCD_GEMM_DEBUG
extern "C" int tcd_gemm_debug_read(unsigned long long* host8, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(host8, tcd::g_gemm_dbg, 8 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(tcd::g_gemm_dbg, z, sizeof(z));
  }
  return 0;
}
#endif
