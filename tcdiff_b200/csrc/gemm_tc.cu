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
#include <stdlib.h>

#include "tc_gemm_common.cuh"

namespace tcd {

// CONV: converged producer / MMA issue loops (tc_gemm_common.cuh, `_p` wrappers); 0 keeps the lane-0 loops.
// ROWSTORE: unaligned output pitch -> rows transposed through shared memory (epilogue_drain); float / no activation only.
template <typename OutT, int ACT, int CONV, bool ROWSTORE = false>
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
    if constexpr (CONV != 0) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx_p(leader, full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          tma_load_2d_p(leader, sa, &tmap_a, full_bar(stage), kb * BK, m0);
          tma_load_2d_p(leader, sa + A_STAGE_BYTES, &tmap_b, full_bar(stage), kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (lane == 0) {
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
    if constexpr (CONV != 0) {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc_mma_f16_p(leader, tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (uint32_t)(kb | k));
          tc_commit_p(leader, empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_p(leader, tfull_bar(as));
      }
    } else if (lane == 0) {
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
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
      epilogue_drain<OutT, ACT, ROWSTORE>(taddr, row0, n0 + half * (BN / 2), lane, slot, &tmap_c, use_tma_store, bias, bias_vec, act,
                                C, ldc, vec_ok, M, N);
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


template <typename OutT, int ACT, int CONV, bool ROWSTORE>
static int launch_tcr(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                      const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<OutT, ACT, CONV, ROWSTORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) { set_error("gemm_bf16_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tc_kernel<OutT, ACT, CONV, ROWSTORE><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ta, tb, tc, use_tma_store, bias, act, (OutT*)C, ldc, M, N, K);
  return check_launch("gemm_bf16_tc");
}

template <typename OutT, int ACT, int CONV>
static int launch_tcv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                     const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  if constexpr (sizeof(OutT) == 4 && ACT == TCD_ACT_NONE) {
    if (!use_tma_store)
      return launch_tcr<OutT, ACT, CONV, true>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);
  }
  return launch_tcr<OutT, ACT, CONV, false>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);
}

template <typename OutT, int ACT>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int use_tma_store,
                     const float* bias, int act, void* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  return launch_tcv<OutT, ACT, 0>(ta, tb, tc, use_tma_store, bias, act, C, ldc, M, N, K, st);   // lane-0 issue loops (see gemm_tc2.cu)
}

int gemm_bf16_tc2(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act, int out_dtype,
                  void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st);

int gemm_bf16_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act, int out_dtype,
                 void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  TCD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && lda % 8 == 0 && ldw % 8 == 0,
              "tcd_gemm(bf16): A/W base and pitch must be 16-byte aligned (lda=%lld ldw=%lld)", (long long)lda, (long long)ldw);
  TCD_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "tcd_gemm(bf16): dimension too large");
  // big-M GEMMs (every per-token nn.Linear, GELU epilogue included) go to the CTA-pair kernel; tiny ones (conditioning
  // path) stay on one CTA.
  if (M >= 512)
    return gemm_bf16_tc2(A, lda, W, ldw, bias, act, out_dtype, C, ldc, M, N, K, st);
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
