// Fused block tail: a D = 512 projection GEMM (`fc`, `linear2`, `linear3`) with the FiLM + residual + LayerNorm (+ rotary)
// tail of the decoder block in its epilogue.
//
//   y      = A W^T (+ bias)                         A (M,K) bf16, W (512,K) bf16 (nn.Linear weight as stored)
//   z      = LN_in(y)            (optional; SBI_MSA.layer_norm, eps 1e-6, model/model.py:68,106)
//   v      = x_in + (1 + scale) * z + shift         (featurewise_affine + residual, model/model.py:171-173,327,334,339;
//                                                    without film: v = x_in + z, without x_in: v = z)
//   x_out  = v (optional)
//   plain  = LN_next(v) -> bf16,  rot = rotary(plain) -> bf16   (operands of the next block)
//
// i.e. tcd_gemm followed by tcd_film_residual_norm (or tcd_layernorm_rotary) without the round trip of y through HBM and
// without the second kernel.
//
// Round-2 design (the round-1 kernel gave one CTA the whole 128 x 512 fp32 row block = all of tensor memory, so main loop and
// epilogue serialised inside an SM and it lost to the unfused pair): a CLUSTER OF TWO CTAs owns a 128-row tile and splits it
// by COLUMNS.  Each CTA computes 128 x 256 of y (UMMA M = 128, N = 256, its half of W), so its accumulator is 256 tensor-memory
// columns and there are TWO of them: the epilogue of tile i overlaps the MMAs of tile i + 1.  The LayerNorm statistics need
// the full 512-column row: every thread owns 128 columns of one row, computes (sum, centred sum of squares) of its segment
// and publishes them to BOTH CTAs (st.shared::cluster into the peer's exchange buffer + mbarrier arrive with cluster-scope
// release); the four segments of a row are merged with the exact pairwise (Chan) formula in a fixed order, so both CTAs
// get identical statistics.  Two exchanges per tile (LN_in, LN_next).
//
// Per CTA (11 warps, one CTA per SM):
//   warp 0      TMA producer, operands: per 64-wide k block A 128 x 64 and W 256 x 64 bf16 (128B swizzle), 2-stage 48 KB ring
//   warp 1      MMA issuer (converged loop, elected lane), tcgen05.commit per stage / per tile
//   warp 2      stages bias / LayerNorm vectors of this CTA's 256 columns in shared memory (broadcast reads afterwards), then
//               TMA producer of the residual: x_in in 128-row x 32-column fp32 boxes (16 KB, 128B swizzle), 3-slot ring; box i
//               of a tile is read by the four epilogue warps of column half (i & 1)
//   warps 3..10 epilogue: thread = (row, 128-column segment); TMEM lane quarter = warp & 3, column half = (warp - 3) >> 2.
//               Tensor-memory reads are the scarce resource (64 B/clk per SM: 2 048 cycles per sweep of the 128 x 256 block),
//               so every statistic is one-pass (sums of (v - c) and (v - c)^2 about an element c of the data) and chunk k + 1
//               is in flight while chunk k is processed; the arithmetic runs two columns per instruction (fma.rn.f32x2).
//               pass 1  LN_in statistics of y (one sweep), exchange
//               pass 2  per 32-column chunk: v = x + (1 + scale) z + shift replaces y in tensor memory (tcgen05.st) and its
//                       statistics accumulate; the x slot is released as soon as the warp has read it; v staged in one of
//                       the warp's two 4 KB buffers, TMA store to x_out; exchange
//               pass 3  per 32 columns: normalise, (rotate with cos / sin staged through the warp's scratch buffer; the next
//                       piece's table rows are fetched while this one is computed,) bf16, stage; one TMA store per 64 columns
// Row tails (M % 128) are zero-filled on load and clipped on store by the tensor maps.
#include <stdlib.h>

#include "tc_gemm_common.cuh"

namespace tcd {

int num_sms();
int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32);

namespace gf {

constexpr int FN = 512;                                  // the full row
constexpr int CN = 256;                                  // columns per CTA
constexpr int SEG = 128;                                 // columns per epilogue thread
constexpr int THREADS = 11 * 32;
constexpr int OA_BYTES = BM * BK * 2;                    // 16 KB
constexpr int OW_BYTES = CN * BK * 2;                    // 32 KB
constexpr int OSTAGE = OA_BYTES + OW_BYTES;              // 48 KB
constexpr int XSLOT = BM * 128;                          // 16 KB: 128 rows x 32 fp32
constexpr int WSLOT = 32 * 128;                          // per-warp buffer: 32 rows x 128 bytes
constexpr int XCH_BYTES = BM * 4 * 8;                    // [128 rows][4 segments] (mean, m2)
constexpr int PAR_BYTES = 5 * CN * 4;                    // bias | gin | bin | gnext | bnext of this CTA's columns
constexpr int FILM_BYTES = 2 * 2 * CN * 4;               // [sample 0 / 1][scale | shift][256 columns]
// Shared-memory plan (227 KB).  DEEP = the K = 1024 feed-forward tail: its main loop is bound by the latency of the operand
// loads (r02: the MMA warp waited 17 500 of 20 400 cycles per tile with two stages in flight), it writes no x_out and one
// output, so it trades the second per-warp buffer and one residual slot for a third operand stage.
template <bool DEEP>
struct Lay {
  static constexpr int OSTAGES = DEEP ? 3 : 2;
  static constexpr int XSLOTS = DEEP ? 2 : 3;
  static constexpr int NBUF = DEEP ? 1 : 2;                        // per-warp output buffers (stg [, scr])
  static constexpr int OFF_X = OSTAGES * OSTAGE;
  static constexpr int OFF_STG = OFF_X + XSLOTS * XSLOT;
  static constexpr int OFF_SCR = OFF_STG + EPI_WARPS * WSLOT;      // only with NBUF == 2
  static constexpr int OFF_XCH = OFF_STG + NBUF * EPI_WARPS * WSLOT;
  static constexpr int OFF_PAR = OFF_XCH + 2 * XCH_BYTES;
  static constexpr int OFF_BAR = OFF_PAR + PAR_BYTES;
  static constexpr int OFF_FILM = OFF_BAR + 256;                   // FiLM rows of the tile's (at most two) samples
  static constexpr size_t SMEM = OFF_FILM + FILM_BYTES;            // 230 656 B (both plans) <= 232 448
  // barriers: OSTAGES full, OSTAGES empty, 2 tfull, 2 tempty, XSLOTS xfull, XSLOTS xempty, 8 exchange [buffer][quarter]
  static constexpr int B_EMPTY = OSTAGES, B_TFULL = 2 * OSTAGES, B_TEMPTY = B_TFULL + 2, B_XFULL = B_TEMPTY + 2,
                       B_XEMPTY = B_XFULL + XSLOTS, B_XCH = B_XEMPTY + XSLOTS, NBAR = B_XCH + 8;
  static_assert(8 * NBAR + 4 <= 256 && SMEM <= 232448, "shared-memory plan");
};
constexpr uint32_t XCH_TX = 2 * 32 * 8;                  // per exchange and lane quarter: the peer's two warps x 32 rows x 8 bytes

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// Remote (or own-CTA) shared-memory store whose completion is counted on an mbarrier of the destination CTA: the reader only
// waits on that barrier (no cluster-scope release / acquire: those compile to MEMBAR + ERRBAR on the arrive and CCTL.IVALL — an
// invalidation of the whole L1 — on every wait; ncu r02: 12 % of the warp stalls were membar and the FiLM rows missed L1 after
// every exchange)
__device__ __forceinline__ void st_async_f2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_expect_only(uint32_t bar, uint32_t bytes) {   // raises the byte expectation, no arrival
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// tcgen05.ld and its wait in ONE asm statement.  As separate statements (load chunk k + 1, compute chunk k, then wait) the
// compiler is free to move or spill the destination registers between the two — it does not know the load is asynchronous —
// and under this kernel's register pressure it did: r02, the last four columns of a chunk were occasionally stale.
__device__ __forceinline__ void tc_ld32_sync(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// (cp.async.bulk.prefetch.tensor of the next tile's A / x boxes into L2 was measured twice in r02 and made every tail 15-20 %
// slower: the prefetches compete with the demand loads for the same L2 -> SM path.)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool DBG>
__device__ __forceinline__ long long tick() {
  if constexpr (DBG) return clock64();
  else return 0;
}

struct Params {
  const float* bias;        // (512) or NULL
  const float* gin;         // LN_in gamma / beta or NULL
  const float* bin;
  float eps_in;
  const float* film;        // (samples, film_ld): [scale(512) | shift(512)] at film_off, or NULL
  int64_t film_ld, film_off;
  const float* gnext;       // LN_next gamma / beta
  const float* bnext;
  float eps_next;
  const float* rot_cos;     // transposed tables (256 angles, rot_ld >= tokens_per_sample positions) or NULL
  const float* rot_sin;
  int64_t rot_ld;
  int has_xin, has_xout, has_plain, has_rot;
  int M, K, tps;
  unsigned long long* dbg;  // optional 8 cycle counters (tcd_gemm_frn_set_debug): epilogue warp 3 lane 0 / MMA warp of every CTA
};

// F: which optional parts of the tail exist, as compile-time constants for the three decoder tails (with run-time flags every
// 4-column step of pass 2 became its own basic block and each shared / global load was consumed by the next instruction:
// ncu r02, half of all warp stalls were the long scoreboard in pass 2); F_GENERIC keeps the run-time flags for any other mix.
constexpr int F_BIAS = 1, F_IN = 2, F_FILM = 4, F_XIN = 8, F_XOUT = 16, F_GENERIC = 32;
// F_NOAFF: LN_next without its affine — the caller folded gamma / beta into the weights of the projection that consumes the
// plain output (W diag(gamma), b + W beta), so pass 3 neither reads the two per-column vectors nor multiplies by them.
constexpr int F_NOAFF = 64;

template <bool DBG, int F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) gemm_frn_kernel(
    const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
    const __grid_constant__ CUtensorMap tm_xin, const __grid_constant__ CUtensorMap tm_xout,
    const __grid_constant__ CUtensorMap tm_plain, const __grid_constant__ CUtensorMap tm_rot, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();                    // SWIZZLE_128B atoms need 1024-byte alignment
  using L = Lay<(F & ~F_NOAFF) == (F_BIAS | F_FILM | F_XIN)>;         // the feed-forward tail gets the deep operand ring
  constexpr int OSTAGES = L::OSTAGES, XSLOTS = L::XSLOTS, OFF_X = L::OFF_X, OFF_STG = L::OFF_STG, OFF_XCH = L::OFF_XCH,
                OFF_PAR = L::OFF_PAR, OFF_BAR = L::OFF_BAR, OFF_FILM = L::OFF_FILM, NBAR = L::NBAR;
  auto bar = [&](int i) { return base + OFF_BAR + 8u * (uint32_t)i; };
  auto full_bar = [&](int s) { return bar(s); };
  auto empty_bar = [&](int s) { return bar(L::B_EMPTY + s); };
  auto tfull_bar = [&](int s) { return bar(L::B_TFULL + s); };
  auto tempty_bar = [&](int s) { return bar(L::B_TEMPTY + s); };
  auto xfull_bar = [&](int s) { return bar(L::B_XFULL + s); };
  auto xempty_bar = [&](int s) { return bar(L::B_XEMPTY + s); };
  auto xch_bar = [&](int b, int q) { return bar(L::B_XCH + b * 4 + q); };
  const uint32_t tmem_slot = bar(NBAR);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + 8 * NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int M = p.M;
  const int num_tiles = (M + BM - 1) / BM;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    if (p.has_xin) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_xin) : "memory");
    for (int s = 0; s < OSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    for (int s = 0; s < XSLOTS; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), EPI_WARPS); }
    for (int b = 0; b < 2; ++b)
      for (int q = 0; q < 4; ++q) { mbar_init(xch_bar(b, q), 64); mbar_expect_only(xch_bar(b, q), XCH_TX); }   // 64 local arrives + the peer's bytes
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2) {
    float* par = reinterpret_cast<float*>(smem_raw + OFF_PAR);
    const float* src[5] = {p.bias, p.gin, p.bin, p.gnext, p.bnext};
#pragma unroll
    for (int v = 0; v < 5; ++v)
      for (int c = lane; c < CN; c += 32) par[v * CN + c] = src[v] != nullptr ? __ldg(src[v] + (int)rank * CN + c) : 0.f;
  }
  tc_fence_before();
  cluster_sync_all();        // barrier inits (the peer arrives on ours), TMEM allocation and parameter vectors are visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_sync();                // the prologue (weights only) overlaps the previous kernel's tail; activations from here on

  if (warp == 0) {
    // ===================== TMA producer: operands =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int m0 = tile * BM;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx_p(leader, full_bar(stage), OSTAGE);
        const uint32_t sa = base + stage * OSTAGE;
        tma_load_2d_p(leader, sa, &tm_a, full_bar(stage), kb * BK, m0);
        tma_load_2d_p(leader, sa + OA_BYTES, &tm_w, full_bar(stage), kb * BK, (int)rank * CN);
        if (++stage == OSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    long long d_tempty = 0, d_loop = 0, d_full = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const long long c0 = tick<DBG>();
      mbar_wait(tempty_bar(as), aphase ^ 1u);               // the epilogue has drained this accumulator
      tc_fence_after();
      const long long c1 = tick<DBG>();
      d_tempty += c1 - c0;
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * CN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const long long f0 = tick<DBG>();
        mbar_wait(full_bar(stage), phase);
        d_full += tick<DBG>() - f0;
        tc_fence_after();
        const uint32_t sa = base + stage * OSTAGE;
        const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + OA_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k)
          tc_mma_f16_p(leader, tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc, (uint32_t)(kb | k));
        tc_commit_p(leader, empty_bar(stage));
        if (++stage == OSTAGES) { stage = 0; phase ^= 1u; }
      }
      tc_commit_p(leader, tfull_bar(as));
      d_loop += tick<DBG>() - c1;
    }
    if (DBG && p.dbg != nullptr && lane == 0) atomicAdd(p.dbg + 7, (unsigned long long)d_tempty);
    (void)d_loop; (void)d_full;
  } else if (warp == 2) {
    // ===================== TMA producer: residual =====================
    if (p.has_xin) {
      const uint32_t leader = elect_one();
      int slot = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m0 = tile * BM;
        for (int i = 0; i < 8; ++i) {                         // box i: column half (i & 1), chunk (i >> 1) of that half
          mbar_wait(xempty_bar(slot), phase ^ 1u);
          mbar_expect_tx_p(leader, xfull_bar(slot), XSLOT);
          tma_load_2d_p(leader, base + OFF_X + slot * XSLOT, &tm_xin, xfull_bar(slot),
                        (int)rank * CN + (i & 1) * SEG + (i >> 1) * 32, m0);
          if (++slot == XSLOTS) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 3) {
    // ===================== epilogue (8 warps; thread = one row x 128 columns) =====================
    const int ew = warp - 3;
    const int quarter = warp & 3;                          // TMEM lanes [32*quarter, +32) are visible to this warp
    const int half = ew >> 2;
    const int r = quarter * 32 + lane;                     // row inside the tile
    const int seg = (int)rank * 2 + half;                  // which 128 columns of the 512-column row
    const int ccol = half * SEG;                           // first column inside this CTA's 256
    const int gcol = seg * SEG;                            // first global column
    const uint32_t peer = rank ^ 1u;
    constexpr bool two_bufs = L::NBUF == 2;              // one buffer: every reuse waits for the previous store's read
    const uint32_t stg = base + OFF_STG + (uint32_t)(ew * WSLOT), scr = two_bufs ? base + L::OFF_SCR + (uint32_t)(ew * WSLOT) : stg;
    const uint32_t lane_t = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ccol;
    const bool generic = (F & F_GENERIC) != 0;
    const bool has_bias = generic ? p.bias != nullptr : (F & F_BIAS) != 0, has_in = generic ? p.gin != nullptr : (F & F_IN) != 0;
    const bool has_film = generic ? p.film != nullptr : (F & F_FILM) != 0, has_xin = generic ? p.has_xin != 0 : (F & F_XIN) != 0;
    const bool has_xout = generic ? p.has_xout != 0 : (F & F_XOUT) != 0;
    constexpr bool no_aff = (F & F_GENERIC) == 0 && (F & F_NOAFF) != 0;
    // plain (schedulable) views of shared memory: parameter vectors of this thread's 128 columns
    const uint8_t* xring = smem_raw + OFF_X + r * 128;
    const float4* parv = reinterpret_cast<const float4*>(smem_raw + OFF_PAR + ccol * 4);   // + v * (CN / 4): bias, gin, bin, gnext, bnext
    long long d_xw = 0, d_ex = 0;
    uint32_t xk = 0;                                       // running exchange count (buffer / barrier = xk & 1)
    constexpr float invS = 1.0f / (float)SEG, invD = 1.0f / (float)FN;

    // (mean, centred sum of squares) of this thread's segment -> mean, rstd of the whole row; the four segments are merged
    // with the exact pairwise formula in a fixed order, so all four threads of the row (two per CTA) get identical values
    auto exchange = [&](float m, float m2, float eps, float& mean, float& rstd) {
      const uint32_t b = xk & 1u, parity = (xk >> 1) & 1u;
      ++xk;
      const uint32_t rowbuf = base + OFF_XCH + b * XCH_BYTES + (uint32_t)(r * 32);
      const uint32_t mine = rowbuf + (uint32_t)(seg * 8);
      const uint32_t xb = xch_bar((int)b, quarter);
      st_shared_f2(mine, m, m2);                                       // own CTA: plain store + arrive (release, CTA scope)
      mbar_arrive(xb);
      st_async_f2(mapa(mine, peer), m, m2, mapa(xb, peer));            // peer CTA: bytes counted on its barrier
      { const long long t0 = tick<DBG>(); mbar_wait(xb, parity); d_ex += tick<DBG>() - t0; }
      // expect the peer's bytes of this barrier's next use (exchange xk + 1) right away: they can only be sent after every
      // warp has passed the exchange in between, so the byte count never runs ahead of the expectation
      if (half == 0 && lane == 0) mbar_expect_only(xb, XCH_TX);
      const float4 a = lds128(rowbuf), c = lds128(rowbuf + 16);       // (m0, q0, m1, q1), (m2, q2, m3, q3)
      mean = ((a.x + a.z) + (c.x + c.z)) * 0.25f;
      const float d0 = a.x - mean, d1 = a.z - mean, d2 = c.x - mean, d3 = c.z - mean;
      const float m2t = ((a.y + a.w) + (c.y + c.w)) + (float)SEG * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
      rstd = 1.0f / sqrtf(m2t * invD + eps);
    };
    // one-pass statistics about a shift c (an element of the data): s2 += (v - c), q2 += (v - c)^2, two columns per lane of
    // the packed fp32 pipe
    auto acc_stats = [&](float2 v, float2 nc, float2& s2, float2& q2) {
      const float2 d = __fadd2_rn(v, nc);
      s2 = __fadd2_rn(s2, d);
      q2 = __ffma2_rn(d, d, q2);
    };
    // cos / sin of piece n (32 columns = 16 angles) for this thread's row from the TRANSPOSED tables (angle-major, position
    // contiguous): lanes hold consecutive positions, so each of the 32 loads is one coalesced 128-byte request — no staging
    // through shared memory (r02: the row-major tables went global -> registers -> shared -> registers, 3 x the traffic
    // through the shared-memory pipe and two warp barriers per piece)
    auto rot_fetch = [&](int pr, int n, float (&cs)[16], float (&sn)[16]) {
      const float* ct = p.rot_cos + (int64_t)((gcol + 32 * n) / 2) * p.rot_ld + pr;
      const float* st = p.rot_sin + (int64_t)((gcol + 32 * n) / 2) * p.rot_ld + pr;
#pragma unroll
      for (int i = 0; i < 16; ++i) { cs[i] = __ldg(ct + (int64_t)i * p.rot_ld); sn[i] = __ldg(st + (int64_t)i * p.rot_ld); }
    };

    int it = 0;
    long long d_e[5] = {0, 0, 0, 0, 0};                    // cycles: wait for the MMAs, pass 1, 2, exchange 2, pass 3
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const uint32_t acc = lane_t + (uint32_t)(as * CN);
      const int m0 = tile * BM;
      const int row0 = m0 + quarter * 32;
      const bool live = row0 < M;                          // warp-uniform: some row of this warp exists
      const int grow = min(m0 + r, M - 1);                 // clamped: rows past M compute garbage that is clipped on store
      // FiLM rows: a 128-row tile touches at most two samples when a sample has at least 128 tokens; their scale | shift rows
      // (this CTA's 256 columns) are staged in shared memory by the eight epilogue warps, one 16-byte piece per thread, before
      // the wait for the MMAs.  (As global loads they missed the small L1 left beside 226 KB of shared memory behind the
      // rotary-table stream and every chunk of pass 2 paid an L2 round trip: ncu r02, long-scoreboard stalls on `scale + 1`.)
      const bool film_smem = has_film && p.tps >= BM;
      const int samp0 = m0 / p.tps;
      const float* fp = (has_film && !film_smem) ? p.film + (int64_t)(grow / p.tps) * p.film_ld + p.film_off + gcol : nullptr;
      const float4* fs = reinterpret_cast<const float4*>(smem_raw + OFF_FILM) + ((grow / p.tps - samp0) & 1) * (2 * CN / 4) + ccol / 4;
      if (film_smem) {
        asm volatile("bar.sync 1, 256;" ::: "memory");    // every epilogue warp is past pass 2 of the previous tile
        const int idx = ew * 32 + lane;                    // [sample 2][scale | shift 2][64 pieces]
        const int smp = min(samp0 + (idx >> 7), (M - 1) / p.tps);
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.film + (int64_t)smp * p.film_ld + p.film_off + ((idx >> 6) & 1) * FN +
                                                                 (int)rank * CN) + (idx & 63));
        reinterpret_cast<float4*>(smem_raw + OFF_FILM)[idx] = v4;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const long long e0 = tick<DBG>();
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const long long e1 = tick<DBG>();

      uint32_t cur[32];                                    // one 32-column chunk of this thread's row

      // ---- pass 1: statistics of y for the inner LayerNorm (one sweep)
      float mean1 = 0.f, rstd1 = 1.f;
      if (has_in) {
        float c = 0.f;
        float2 nc = make_float2(0.f, 0.f);
        float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_ld32_sync(acc + (uint32_t)(32 * k), cur);
          if (k == 0) { c = __uint_as_float(cur[0]); nc = make_float2(-c, -c); }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bq = parv[8 * k + j];
              acc_stats(__fadd2_rn(make_float2(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1])), make_float2(bq.x, bq.y)), nc, s2, q2);
              acc_stats(__fadd2_rn(make_float2(__uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3])), make_float2(bq.z, bq.w)), nc, s2, q2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              acc_stats(make_float2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), nc, s2, q2);
          }
        }
        const float s = s2.x + s2.y, q = q2.x + q2.y;
        exchange(c + s * invS, fmaxf(q - s * s * invS, 0.f), p.eps_in, mean1, rstd1);
      }

      const long long e2 = tick<DBG>();
      // ---- pass 2: v = x + (1 + scale) * z + shift, chunk by chunk, into tensor memory (replacing y) and to x_out.
      //      Staging alternates between the warp's two buffers; whatever pass 3 of the previous tile stored from them
      //      has long been read, but the bulk-group accounting needs the explicit wait.
      if (has_xout) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      const float2 r1 = make_float2(rstd1, rstd1), nm1 = make_float2(-mean1 * rstd1, -mean1 * rstd1);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f), ncv = make_float2(0.f, 0.f);
      float cv = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        tc_ld32_sync(acc + (uint32_t)(32 * k), cur);
        uint32_t xslot = 0;
        if (has_xin) {
          const uint32_t xseq = (uint32_t)it * 8u + (uint32_t)(2 * k + half);
          xslot = xseq % (uint32_t)XSLOTS;
          // A slot alternates between boxes of both column halves (3 slots, 2 halves), and a parity wait can only tell a
          // barrier's current phase from the previous one: a warp that skipped the other half's fills could find its slot's
          // barrier two phases on and read a box that has not landed (r02: 16-byte pieces of stale x, about once per ten
          // launches).  So all eight warps pass every box in order — wait for the fill, arrive on the release barrier — and
          // only the owners read it in between.
          const long long t0 = tick<DBG>();
          if (half == 1) {
            const uint32_t o = xseq - 1u, os = o % (uint32_t)XSLOTS;
            mbar_wait(xfull_bar((int)os), (o / (uint32_t)XSLOTS) & 1u);
            if (lane == 0) mbar_arrive(xempty_bar((int)os));
          }
          mbar_wait(xfull_bar((int)xslot), (xseq / (uint32_t)XSLOTS) & 1u);
          d_xw += tick<DBG>() - t0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float2 ya = make_float2(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]));
          float2 yb = make_float2(__uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3]));
          if (has_bias) {
            const float4 bq = parv[8 * k + j];
            ya = __fadd2_rn(ya, make_float2(bq.x, bq.y));
            yb = __fadd2_rn(yb, make_float2(bq.z, bq.w));
          }
          if (has_in) {
            const float4 g = parv[CN / 4 + 8 * k + j];
            const float4 bb = parv[2 * (CN / 4) + 8 * k + j];
            ya = __ffma2_rn(__ffma2_rn(ya, r1, nm1), make_float2(g.x, g.y), make_float2(bb.x, bb.y));
            yb = __ffma2_rn(__ffma2_rn(yb, r1, nm1), make_float2(g.z, g.w), make_float2(bb.z, bb.w));
          }
          if (has_film) {
            const float4 sc = film_smem ? fs[8 * k + j] : __ldg(reinterpret_cast<const float4*>(fp + 32 * k) + j);
            const float4 sh = film_smem ? fs[CN / 4 + 8 * k + j] : __ldg(reinterpret_cast<const float4*>(fp + FN + 32 * k) + j);
            const float2 one = make_float2(1.0f, 1.0f);
            ya = __ffma2_rn(__fadd2_rn(make_float2(sc.x, sc.y), one), ya, make_float2(sh.x, sh.y));
            yb = __ffma2_rn(__fadd2_rn(make_float2(sc.z, sc.w), one), yb, make_float2(sh.z, sh.w));
          }
          if (has_xin) {
            const float4 xv = *reinterpret_cast<const float4*>(xring + xslot * XSLOT + (((j ^ lane) & 7) << 4));   // SWIZZLE_128B: 16-byte chunk j of row r
            ya = __fadd2_rn(make_float2(xv.x, xv.y), ya);
            yb = __fadd2_rn(make_float2(xv.z, xv.w), yb);
          }
          if (k == 0 && j == 0) { cv = ya.x; ncv = make_float2(-cv, -cv); }
          acc_stats(ya, ncv, s2, q2);
          acc_stats(yb, ncv, s2, q2);
          cur[4 * j] = __float_as_uint(ya.x); cur[4 * j + 1] = __float_as_uint(ya.y);
          cur[4 * j + 2] = __float_as_uint(yb.x); cur[4 * j + 3] = __float_as_uint(yb.y);
        }
        tc_st32(acc + (uint32_t)(32 * k), cur);
        if (has_xin) {
          // The box may only be released once its loads have RETURNED, not merely been issued: ptxas batches them right before
          // the arrive and consumes them after it, and a load still queued in the memory pipe when the producer's refill
          // lands reads the NEXT box (r02, decoded with a structured x: the last 16 bytes of a row came from the box that
          // took the slot over).  v depends on every load of the chunk and the tensor-memory store above consumes all of v,
          // so releasing after it is a true dependency the scheduler has to honour.  (Releasing right after the loads,
          // behind a dependent shared-memory store, was measured: the eight float4 held across the chunk spill, +8 %.)
          __syncwarp();                                    // every lane has read its row of the box
          if (lane == 0) mbar_arrive(xempty_bar((int)xslot));
          if (half == 0) {                                 // pass the other half's box of this step
            const uint32_t o = (uint32_t)it * 8u + (uint32_t)(2 * k + 1), os = o % (uint32_t)XSLOTS;
            mbar_wait(xfull_bar((int)os), (o / (uint32_t)XSLOTS) & 1u);
            if (lane == 0) mbar_arrive(xempty_bar((int)os));
          }
        }
        if (has_xout) {
          const uint32_t buf = (k & 1) ? scr : stg;
          if (k >= 2 || !two_bufs) {                       // the store of chunk k-2 has read this buffer (k-1 may be pending)
            if (lane == 0) { if (two_bufs) tma_store_wait_read1(); else tma_store_wait_read(); }
            __syncwarp();
          }
          const uint32_t wb = buf + (uint32_t)(lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(wb + (uint32_t)(((j ^ lane) & 7) << 4), cur[4 * j], cur[4 * j + 1], cur[4 * j + 2], cur[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && live) {
            tma_store_2d(&tm_xout, buf, gcol + 32 * k, row0);
            tma_store_commit();
          }
        }
      }
      const long long e3 = tick<DBG>();

      // ---- statistics of v: exchange; the first rotary table piece is fetched before the wait
      float csa[16], sna[16], csb[16], snb[16];            // rotary factors of the current / next piece
      const int pr = min(row0 + lane, M - 1) % p.tps;      // this row's position inside its sample
      if (p.has_rot) rot_fetch(pr, 0, csa, sna);
      float mean2, rstd2;
      {
        const float s = s2.x + s2.y, q = q2.x + q2.y;
        exchange(cv + s * invS, fmaxf(q - s * s * invS, 0.f), p.eps_next, mean2, rstd2);
      }
      tc_wait_st();                                        // v is in tensor memory
      const long long e4 = tick<DBG>();

      // ---- pass 3: LayerNorm of v -> bf16 operands of the next block, 64 columns (one 128-byte bf16 row) per store.
      //      Staging alternates between the warp's two buffers.
      if (lane == 0) tma_store_wait_read();                // x_out stores of pass 2 have read both buffers
      __syncwarp();
      const float2 r2 = make_float2(rstd2, rstd2), nm2 = make_float2(-mean2 * rstd2, -mean2 * rstd2);
#pragma unroll 1
      for (int mode = 0; mode < 2; ++mode) {               // 0: plain, 1: rot
        if (mode == 0 ? !p.has_plain : !p.has_rot) continue;
#pragma unroll
        for (int n = 0; n < 4; ++n) {                      // 32-column pieces; a store per two pieces
          tc_ld32_sync(acc + (uint32_t)(32 * n), cur);
          const uint32_t obuf = (n & 2) ? scr : stg;
          const uint32_t ob = obuf + (uint32_t)(lane * 128);
          float nv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 na = __ffma2_rn(make_float2(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1])), r2, nm2);
            float2 nb = __ffma2_rn(make_float2(__uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3])), r2, nm2);
            if constexpr (!no_aff) {
              const float4 g = parv[3 * (CN / 4) + 8 * n + j];
              const float4 bb = parv[4 * (CN / 4) + 8 * n + j];
              na = __ffma2_rn(na, make_float2(g.x, g.y), make_float2(bb.x, bb.y));
              nb = __ffma2_rn(nb, make_float2(g.z, g.w), make_float2(bb.z, bb.w));
            }
            nv[4 * j] = na.x; nv[4 * j + 1] = na.y; nv[4 * j + 2] = nb.x; nv[4 * j + 3] = nb.y;
          }
          if (mode == 1) {
            // interleaved pairs (2i, 2i+1) rotate by angle i of the token's table row (rotary_embedding_torch.py:107-113);
            // the next piece's factors fly during this piece's arithmetic
            float (&cs)[16] = (n & 1) ? csb : csa;
            float (&sn)[16] = (n & 1) ? snb : sna;
            if (n + 1 < 4) rot_fetch(pr, n + 1, (n & 1) ? csa : csb, (n & 1) ? sna : snb);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float n0 = nv[2 * i], n1 = nv[2 * i + 1];
              nv[2 * i] = n0 * cs[i] - n1 * sn[i];
              nv[2 * i + 1] = n1 * cs[i] + n0 * sn[i];
            }
          }
          if ((n & 1) == 0) {                              // the store that last read this buffer is done
            if (lane == 0) { if (two_bufs) tma_store_wait_read1(); else tma_store_wait_read(); }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(ob + (uint32_t)((((4 * (n & 1) + j) ^ lane) & 7) << 4), bf2(nv[8 * j], nv[8 * j + 1]),
                   bf2(nv[8 * j + 2], nv[8 * j + 3]), bf2(nv[8 * j + 4], nv[8 * j + 5]), bf2(nv[8 * j + 6], nv[8 * j + 7]));
          if (n & 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && live) {
              tma_store_2d(mode == 0 ? &tm_plain : &tm_rot, obuf, gcol + 32 * (n - 1), row0);
              tma_store_commit();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));          // the MMAs of tile it + 2 may overwrite this accumulator
      d_e[0] += e1 - e0; d_e[1] += e2 - e1; d_e[2] += e3 - e2; d_e[3] += e4 - e3; d_e[4] += tick<DBG>() - e4;
    }
    if (lane == 0) tma_store_wait_all();
    if (DBG && p.dbg != nullptr && ew == 0 && lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) atomicAdd(p.dbg + k, (unsigned long long)d_e[k]);
      atomicAdd(p.dbg + 5, (unsigned long long)d_xw); atomicAdd(p.dbg + 6, (unsigned long long)d_ex);
    }
  }
  tc_fence_before();
  cluster_sync_all();        // no CTA may exit while its peer can still write its exchange buffers / barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace gf

static unsigned long long* g_frn_dbg = nullptr;

int gemm_frn_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int64_t M, int64_t K,
                  const float* x_in, float* x_out, const float* gin, const float* bin, float eps_in, const float* film,
                  int64_t film_ld, int64_t film_off, const float* gnext, const float* bnext, float eps_next, void* out_plain,
                  void* out_rot, const float* rot_cos, const float* rot_sin, int64_t rot_ld, int tps, cudaStream_t st) {
  CUtensorMap ta, tw, txi, txo, tp, tr;
  int rc = make_tmap_2d(&ta, A, M, K, lda, BM, false);
  if (rc) return rc;
  rc = make_tmap_2d(&tw, W, gf::FN, K, ldw, gf::CN, false);
  if (rc) return rc;
  txi = ta;
  if (x_in) { rc = make_tmap_2d(&txi, x_in, M, gf::FN, gf::FN, BM, true); if (rc) return rc; }
  txo = ta;
  if (x_out) { rc = make_tmap_2d(&txo, x_out, M, gf::FN, gf::FN, 32, true); if (rc) return rc; }
  tp = ta;
  if (out_plain) { rc = make_tmap_2d(&tp, out_plain, M, gf::FN, gf::FN, 32, false); if (rc) return rc; }
  tr = ta;
  if (out_rot) { rc = make_tmap_2d(&tr, out_rot, M, gf::FN, gf::FN, 32, false); if (rc) return rc; }
  gf::Params p;
  p.bias = bias; p.gin = gin; p.bin = bin; p.eps_in = eps_in;
  p.film = film; p.film_ld = film_ld; p.film_off = film_off;
  p.gnext = gnext; p.bnext = bnext; p.eps_next = eps_next;
  p.rot_cos = rot_cos; p.rot_sin = rot_sin; p.rot_ld = rot_ld;
  p.has_xin = x_in != nullptr; p.has_xout = x_out != nullptr; p.has_plain = out_plain != nullptr; p.has_rot = out_rot != nullptr;
  p.M = (int)M; p.K = (int)K; p.tps = tps;
  p.dbg = g_frn_dbg;
  const int tiles = (int)((M + BM - 1) / BM);
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  const int flags = (bias ? gf::F_BIAS : 0) | (gin ? gf::F_IN : 0) | (film ? gf::F_FILM : 0) | (x_in ? gf::F_XIN : 0) |
                    (x_out ? gf::F_XOUT : 0) | (gnext ? 0 : gf::F_NOAFF);
  constexpr int F_SA = gf::F_IN | gf::F_FILM | gf::F_XIN | gf::F_XOUT;     // self- / cross-attention tail
  constexpr int F_FF = gf::F_BIAS | gf::F_FILM | gf::F_XIN;                // feed-forward tail (dead residual)
#define TCD_FRN_LAUNCH(DBGV, FV)                                                                                              \
  do {                                                                                                                        \
    static bool configured = false;                                                                                           \
    if (!configured) {                                                                                                        \
      cudaError_t e = cudaFuncSetAttribute(gf::gemm_frn_kernel<DBGV, FV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gf::Lay<false>::SMEM); \
      if (e != cudaSuccess) { set_error("gemm_frn: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }       \
      configured = true;                                                                                                      \
    }                                                                                                                         \
    cudaError_t le = launch_pdl(gf::gemm_frn_kernel<DBGV, FV>, dim3(2 * pairs), dim3(gf::THREADS), gf::Lay<false>::SMEM, st, tiles <= 2 * max_pairs, ta, tw, txi, txo, tp, tr, p); \
    if (le != cudaSuccess) { set_error("gemm_frn: launch: %s", cudaGetErrorString(le)); return TCD_ERR_CUDA; }                \
  } while (0)
  constexpr int F_SA_N = F_SA | gf::F_NOAFF, F_FF_N = F_FF | gf::F_NOAFF;   // ... with LN_next's affine folded downstream
  if ((flags & gf::F_NOAFF) && flags != F_SA_N && flags != F_FF_N) {
    set_error("gemm_frn: LN_next without affine (next_gamma NULL) is built for the attention and feed-forward tails only");
    return TCD_ERR_INVALID;
  }
  if (p.dbg != nullptr) {
    if (flags == F_SA) TCD_FRN_LAUNCH(true, F_SA);
    else if (flags == F_FF) TCD_FRN_LAUNCH(true, F_FF);
    else if (flags == F_SA_N) TCD_FRN_LAUNCH(true, F_SA_N);
    else if (flags == F_FF_N) TCD_FRN_LAUNCH(true, F_FF_N);
    else TCD_FRN_LAUNCH(true, gf::F_GENERIC);
  } else {
    if (flags == F_SA) TCD_FRN_LAUNCH(false, F_SA);
    else if (flags == F_FF) TCD_FRN_LAUNCH(false, F_FF);
    else if (flags == F_SA_N) TCD_FRN_LAUNCH(false, F_SA_N);
    else if (flags == F_FF_N) TCD_FRN_LAUNCH(false, F_FF_N);
    else TCD_FRN_LAUNCH(false, gf::F_GENERIC);
  }
#undef TCD_FRN_LAUNCH
  return check_launch("gemm_frn");
}

}  // namespace tcd

// Profiling aid: buf = 8 device uint64 counters that every later launch of the fused kernel adds its cycle counts to (epilogue
// warp 3 of every CTA: [0] waiting for the MMAs, [1] pass 1, [2] pass 2, [3] second exchange, [4] pass 3, [5] of pass 2: waiting
// for residual boxes, [6] of passes 1-2: waiting in the exchanges; MMA warp: [7] waiting for the epilogue), or NULL to switch
// it off.
extern "C" int tcd_gemm_frn_set_debug(void* buf) {
  tcd::g_frn_dbg = reinterpret_cast<unsigned long long*>(buf);
  return TCD_OK;
}

extern "C" int tcd_gemm_film_residual_norm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                           int64_t M, int64_t K, const float* x_in, float* x_out,
                                           const float* ln_in_gamma, const float* ln_in_beta, float ln_in_eps,
                                           const float* film, int64_t film_ld, int64_t film_off,
                                           const float* next_gamma, const float* next_beta, float next_eps,
                                           void* out_plain, void* out_rot, const float* rot_cos_t, const float* rot_sin_t,
                                           int64_t rot_ld, int tokens_per_sample, void* stream) {
  const float* rot_cos = rot_cos_t;
  const float* rot_sin = rot_sin_t;
  using namespace tcd;
  TCD_REQUIRE(A && W, "tcd_gemm_film_residual_norm: null pointer");
  TCD_REQUIRE((next_gamma == nullptr) == (next_beta == nullptr), "tcd_gemm_film_residual_norm: next LN params");
  TCD_REQUIRE(next_gamma || !out_rot, "tcd_gemm_film_residual_norm: the rotated operand needs LN_next's affine (it does not commute with the rotation)");
  TCD_REQUIRE((ln_in_gamma == nullptr) == (ln_in_beta == nullptr), "tcd_gemm_film_residual_norm: inner LN params");
  TCD_REQUIRE(out_plain || out_rot, "tcd_gemm_film_residual_norm: no output operand requested");
  TCD_REQUIRE(!out_rot || (rot_cos && rot_sin && rot_ld >= tokens_per_sample),
              "tcd_gemm_film_residual_norm: rotary tables missing (transposed: 256 angles x rot_ld >= tokens_per_sample positions)");
  TCD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && lda % 8 == 0 && ldw % 8 == 0,
              "tcd_gemm_film_residual_norm: A/W base and pitch must be 16-byte aligned");
  TCD_REQUIRE(((uintptr_t)x_in % 16 == 0) && ((uintptr_t)x_out % 16 == 0) && ((uintptr_t)out_plain % 16 == 0) &&
                  ((uintptr_t)out_rot % 16 == 0) && ((uintptr_t)film % 16 == 0) && ((uintptr_t)rot_cos % 4 == 0) &&
                  ((uintptr_t)rot_sin % 4 == 0),
              "tcd_gemm_film_residual_norm: 16-byte alignment");
  TCD_REQUIRE(film_ld % 4 == 0 && film_off % 4 == 0, "tcd_gemm_film_residual_norm: film alignment");
  TCD_REQUIRE(tokens_per_sample > 0 && M < (1LL << 31) && K > 0 && K < (1LL << 31), "tcd_gemm_film_residual_norm: bad shape");
  if (M == 0) return TCD_OK;
  return gemm_frn_bf16(A, lda, W, ldw, bias, M, K, x_in, x_out, ln_in_gamma, ln_in_beta, ln_in_eps, film, film_ld, film_off,
                       next_gamma, next_beta, next_eps, out_plain, out_rot, rot_cos, rot_sin, rot_ld, tokens_per_sample,
                       as_stream(stream));
}
