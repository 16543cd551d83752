// Weight-gradient GEMM of the training step on the tcgen05 tensor cores:  C (M,N) fp32 = A^T B
//   A (K,M) bf16, pitch lda — the output gradient dY as stored (rows = tokens)     -> UMMA operand A, MN-major
//   B (K,N) bf16, pitch ldb — the layer input X as stored (rows = tokens)          -> UMMA operand B, MN-major
// i.e. dW = dY^T X for nn.Linear (autograd of model/model.py's Linear layers, TCDiff.py:232) WITHOUT transposing the
// (tokens x features) activations in memory: both operands are read as MN-major tiles.  A shared-memory stage is
// 64 token rows of 128 (A) + 256 (B) features, brought in by TMA as [64 rows x 64 columns] 128B-swizzled boxes
// (8 KiB each, consecutive 64-column blocks 8 KiB apart = the descriptor's leading byte offset; 8-row groups 1 KiB
// apart = the stride byte offset); one tcgen05.mma K-step = 16 token rows = +2 KiB.
//
// The token dimension is the reduction (K ~ 1e5) while M x N is small (<= 1024 x 2560), so the work is split along K:
// work item = (m tile, n tile, split); each item writes an fp32 partial tile to a workspace, and a second tiny kernel
// sums the splits in a fixed order (deterministic; no atomics).  Same persistent warp-specialised structure, TMEM
// double buffering and TMA-store epilogue as gemm_tc.cu.
#include "tc_gemm_common.cuh"

namespace tcd {

int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32);
int num_sms();

constexpr bool kWgradConvDefault = false;

namespace wg {

constexpr int BOX_BYTES = 64 * 128;            // [64 token rows][64 features] bf16

// MN-major, SWIZZLE_128B: LBO = 8 KiB between 64-feature blocks, SBO = 1 KiB between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(BOX_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// D=f32, A=B=bf16, A and B MN-major (bits 15, 16), N=256, M=128
constexpr uint32_t kIdescMN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                              ((uint32_t)(BM >> 4) << 24);

// CONV: converged producer / MMA issue loops (tc_gemm_common.cuh, `_p` wrappers); false keeps the lane-0 loops.
template <bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tn_bf16_kernel(
    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
    const __grid_constant__ CUtensorMap tmap_c, float* __restrict__ ws, int64_t ws_ld, int M, int N, int K, int splits,
    int kb_per_split) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + EPI_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int num_items = tiles_m * tiles_n * splits;          // item -> (split fastest, n tile, m tile)
  const int num_kb = (K + BK - 1) / BK;
  const int m_pad = tiles_m * BM;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if constexpr (CONV) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int sp = item % splits, tile = item / splits;
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx_p(leader, full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < BM / 64; ++i) tma_load_2d_p(leader, sa + i * BOX_BYTES, &tmap_a, full_bar(stage), m0 + 64 * i, kb * BK);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_2d_p(leader, sa + A_STAGE_BYTES + i * BOX_BYTES, &tmap_b, full_bar(stage), n0 + 64 * i, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int sp = item % splits, tile = item / splits;
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * BOX_BYTES, &tmap_a, full_bar(stage), m0 + 64 * i, kb * BK);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_2d(sa + A_STAGE_BYTES + i * BOX_BYTES, &tmap_b, full_bar(stage), n0 + 64 * i, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if constexpr (CONV) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int sp = item % splits;
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc_mma_f16_p(leader, tmem_d, umma_desc_mn128(sa + k * 2048), umma_desc_mn128(sa + A_STAGE_BYTES + k * 2048), kIdescMN,
                         (kb > kb0 || k != 0) ? 1u : 0u);
          tc_commit_p(leader, empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_p(leader, tfull_bar(as));
      }
    } else if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int sp = item % splits;
        const int kb0 = sp * kb_per_split, kb1 = min(num_kb, kb0 + kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)        // 16 token rows = 2 KiB further into every 8 KiB box
            tc_mma_f16(tmem_d, umma_desc_mn128(sa + k * 2048), umma_desc_mn128(sa + A_STAGE_BYTES + k * 2048), kIdescMN,
                       (kb > kb0 || k != 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));
      }
    }
  } else {
    // ===================== epilogue (8 warps): fp32 partial tile -> workspace[split] =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const uint32_t slot = epi_base + (uint32_t)(ew * EPI_SLOT_BYTES);
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int sp = item % splits, tile = item / splits;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row0 = sp * m_pad + m0 + quarter * 32;       // row in the (splits * m_pad, N) workspace
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
      epilogue_drain<float, TCD_ACT_NONE>(taddr, row0, n0 + half * (BN / 2), lane, slot, &tmap_c, 1, nullptr, false, 0, ws,
                                          ws_ld, true, splits * m_pad, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// C[m, n] = sum over splits of ws[s][m][n], fixed order
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int64_t ws_ld, int64_t split_stride,
                                                            int splits, float* __restrict__ C, int64_t ldc, int M, int N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  const float* p = ws + (int64_t)m * ws_ld + n;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += __ldg(p + s * split_stride);
  C[(int64_t)m * ldc + n] = acc;
}

static void plan(int64_t M, int64_t N, int64_t K, int* splits, int* kb_per_split) {
  const int tiles = (int)(((M + BM - 1) / BM) * ((N + BN - 1) / BN));
  const int num_kb = (int)((K + BK - 1) / BK);
  int s = num_sms() / tiles;                    // fill the machine once
  const int max_s = num_kb / 4 > 0 ? num_kb / 4 : 1;     // at least 4 k-blocks per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  int per = (num_kb + s - 1) / s;
  s = (num_kb + per - 1) / per;                 // no empty splits
  *splits = s;
  *kb_per_split = per;
}

}  // namespace wg
}  // namespace tcd

using namespace tcd;

extern "C" int64_t tcd_gemm_tn_workspace_floats(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int splits, per;
  wg::plan(M, N, K, &splits, &per);
  const int64_t m_pad = (M + BM - 1) / BM * BM, ld = (N + 3) / 4 * 4;
  return (int64_t)splits * m_pad * ld;
}

extern "C" int tcd_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                           int64_t N, int64_t K, float* workspace, void* stream) {
  TCD_REQUIRE(M >= 0 && N >= 0 && K >= 0, "tcd_gemm_tn: bad shape");
  if (M == 0 || N == 0) return TCD_OK;
  TCD_REQUIRE(A && B && C && workspace, "tcd_gemm_tn: null pointer");
  TCD_REQUIRE(K > 0, "tcd_gemm_tn: empty reduction");
  TCD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && lda % 8 == 0 && ldb % 8 == 0 &&
              (uintptr_t)workspace % 16 == 0, "tcd_gemm_tn: bases and pitches must be 16-byte aligned");
  TCD_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "tcd_gemm_tn: dimension too large");
  cudaStream_t st = as_stream(stream);
  int splits, per;
  wg::plan(M, N, K, &splits, &per);
  const int64_t m_pad = (M + BM - 1) / BM * BM, ws_ld = (N + 3) / 4 * 4;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_2d(&ta, A, K, M, lda, 64, false))) return rc;     // (token rows, features): box 64 x 64
  if ((rc = make_tmap_2d(&tb, B, K, N, ldb, 64, false))) return rc;
  if ((rc = make_tmap_2d(&tc, workspace, splits * m_pad, N, ws_ld, 32, true))) return rc;
  constexpr bool conv = kWgradConvDefault;             // lane-0 issue loops (r01 A/B: no difference for the wgrad GEMM)
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wg::gemm_tn_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wg::gemm_tn_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) { set_error("gemm_tn: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int64_t items = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  if (conv)
    wg::gemm_tn_bf16_kernel<true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ta, tb, tc, workspace, ws_ld, (int)M, (int)N, (int)K, splits, per);
  else
    wg::gemm_tn_bf16_kernel<false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ta, tb, tc, workspace, ws_ld, (int)M, (int)N, (int)K, splits, per);
  if ((rc = check_launch("gemm_tn"))) return rc;
  const int64_t total = M * N;
  wg::splitk_reduce_kernel<<<(unsigned)ceil_div(total, (int64_t)256), 256, 0, st>>>(workspace, ws_ld, m_pad * ws_ld, splits, C, ldc,
                                                                                  (int)M, (int)N);
  return check_launch("splitk_reduce");
}
