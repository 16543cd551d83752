// bf16 attention BACKWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 64, no mask.
// Gradients of O = softmax(scale Q K^T) V per (sample, head) — the reverse pass of SBI_MSA's core
// (model/model.py:97-102) and of the music encoder's nn.MultiheadAttention core in the training step
// (model/diffusion.py:636-753, TCDiff.py:232).
//
// Flash-style: nothing of size Lq x Lk is stored; P is recomputed from the forward's log-sum-exp,
//     P = exp2(c s - lse),   dP = dO V^T,   dS = P o (dP - D),   D = rowsum(dO o O),
//     dV = P^T dO,           dK = scale dS^T Q,                  dQ = scale dS K.
// Two deterministic kernels from ONE template (no atomics; S and dP are recomputed in both):
//   DKDV : one CTA work item = 128 keys of one (sample, head); K, V tiles resident; Q, dO streamed 64 rows at a time.
//          TMEM lanes = keys:   T1 = K Q^T (= S^T), T2 = V dO^T (= dP^T);  dV += P^T dO,  dK += dS^T Q.
//   DQ   : one CTA work item = 128 queries; Q, dO tiles resident; K, V streamed 64 keys at a time.
//          TMEM lanes = queries: T1 = Q K^T (= S),  T2 = dO V^T (= dP);     dQ += dS K.
// In both, the second-stage B operand is the streamed tile itself read as an MN-major operand (the tile that was the
// K-major B operand of the first stage), so nothing is ever transposed in memory.
// Roles per CTA (2 CTAs per SM: <= 98 KiB smem, 256 TMEM columns): warp 0 TMA producer, warp 1 MMA issuer,
// warps 2..9 recompute/epilogue with two threads per TMEM lane (32 columns each).
// TMEM: T1 [0,64) | T2 [64,128) | ACC1 [128,192) | ACC2 [192,256).
// Rows past the sequence ends are zero-filled by TMA on load (their contributions vanish: zero K/V rows give zero
// products, zero Q/dO rows with (lse, D) = (0, 0) give dS = 0) and clipped by TMA on store.
#include "tc_attn_common.cuh"
#include "dropout.cuh"

namespace tcd {
constexpr bool kTrainConvDefault = true;    // r01 A/B: dK/dV+dQ 1.166 -> 0.990 ms (self, dropout)

namespace fab {

using namespace fa;

constexpr int BR = 128, BS = 64, HD = 64;      // resident rows (TMEM lanes), streamed rows (TMEM columns), head dim
constexpr int RES_BYTES = BR * 128;            // 128 rows x 64 bf16
constexpr int STR_BYTES = BS * 128;            // 64 rows x 64 bf16
constexpr int SLOT_BYTES = 2 * STR_BYTES;      // (Q, dO) or (K, V)
constexpr int PD_BYTES = BR * 128;             // 128 rows x 64 bf16 : P / dS operand tiles, output staging
constexpr int SM_WARPS = 8;
constexpr int THREADS = (2 + SM_WARPS) * 32;
constexpr int TMEM_COLS = 256;
constexpr int T1_COL = 0, T2_COL = 64, ACC1_COL = 128, ACC2_COL = 192;

template <bool DKDV>
struct Cfg {
  static constexpr int NSLOT = DKDV ? 2 : 3;
  static constexpr int OFF_RES1 = 0, OFF_RES2 = RES_BYTES, OFF_RING = 2 * RES_BYTES;
  static constexpr int OFF_DS = OFF_RING + NSLOT * SLOT_BYTES;
  static constexpr int OFF_P = OFF_DS + PD_BYTES;                       // DKDV only
  static constexpr int OFF_STAT = OFF_P + (DKDV ? PD_BYTES : 0);        // DKDV only: [2][64] (lse, D)
  static constexpr int OFF_BAR = OFF_STAT + (DKDV ? 2 * BS * 8 : 0);
  static constexpr size_t SMEM = 1024 + OFF_BAR + 256;
};

// one thread's 16 columns: p = exp2(c t1 - lse), ds = p (t2 m - D) with m = attention-dropout mask / (1-p) (1 without
// dropout); bf16 into the swizzled operand tile(s): P m (the dV operand) and dS.
// mask coordinates: hash = drop_mix(fixed ^ (var0 + column) * varC) — fixed carries the thread's own (lane) coordinate.
template <bool DKDV, bool DROP>
__device__ __forceinline__ void recompute16(const uint32_t (&t1)[16], const uint32_t (&t2)[16], float c, const float2* cst,
                                            float lse_r, float d_r, uint32_t prow, uint32_t dsrow, int chunk0, int r,
                                            uint32_t fixed, uint32_t var0, uint32_t varC, uint32_t thr, float rk) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float p[8], ds[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float lse, dd;
      if constexpr (DKDV) {
        const float2 s = cst[8 * j + e];           // broadcast read: every lane of the warp reads the same column
        lse = s.x; dd = s.y;
      } else {
        lse = lse_r; dd = d_r;
      }
      const float pe = ex2(fmaf(__uint_as_float(t1[8 * j + e]), c, -lse));
      float m = 1.0f;
      if constexpr (DROP) m = drop_mix(fixed ^ (var0 + (uint32_t)(8 * j + e)) * varC) >= thr ? rk : 0.f;
      ds[e] = pe * (__uint_as_float(t2[8 * j + e]) * m - dd);
      p[e] = pe * m;
    }
    const uint32_t sw = (uint32_t)((((chunk0 + j) ^ r) & 7) << 4);
    if constexpr (DKDV) sts128(prow + sw, pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]));
    sts128(dsrow + sw, pack2(ds[0], ds[1]), pack2(ds[2], ds[3]), pack2(ds[4], ds[5]), pack2(ds[6], ds[7]));
  }
}

// 32 accumulator columns of this thread -> (x mul) -> bf16 -> swizzled staging row
__device__ __forceinline__ void stage32(uint32_t taddr, float mul, uint32_t rowaddr, int chunk0, int r) {
  uint32_t v[32];
  tc_ld32(taddr, v);
  tc_wait_ld();
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128(rowaddr + (uint32_t)((((chunk0 + j) ^ r) & 7) << 4),
           pack2(__uint_as_float(v[8 * j]) * mul, __uint_as_float(v[8 * j + 1]) * mul),
           pack2(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul),
           pack2(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul),
           pack2(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul));
}

// DKDV: res1 = K, res2 = V (box 128), str1 = Q, str2 = dO (box 64), out1 = dV, out2 = dK; Lres = Lk, Lstr = Lq
// DQ  : res1 = Q, res2 = dO (box 128), str1 = K, str2 = V (box 64), out1 = dQ;            Lres = Lq, Lstr = Lk
// stats: (lse, D) per (sample, head, query) as float2
// CONV: the MMA issue loop runs converged (all lanes, one elected lane issues through the `_p` wrappers of
// tc_attn_common.cuh) instead of under `if (lane == 0)`, which ptxas compiles into an R2UR + ELECT loop per UTCHMMA.
template <bool DKDV, bool DROP, bool CONV>
__global__ void __launch_bounds__(THREADS, 2) attention_bwd_tc_kernel(
    const __grid_constant__ CUtensorMap tm_res1, const __grid_constant__ CUtensorMap tm_res2,
    const __grid_constant__ CUtensorMap tm_str1, const __grid_constant__ CUtensorMap tm_str2,
    const __grid_constant__ CUtensorMap tm_out1, const __grid_constant__ CUtensorMap tm_out2,
    const float2* __restrict__ stats, int Lres, int Lstr, int Lq, int heads, int samples, float scale_log2, float scale,
    uint32_t drop_thr, float drop_rk, const uint64_t* __restrict__ rng_state, uint32_t drop_site) {
  using C = Cfg<DKDV>;
  constexpr int NSLOT = C::NSLOT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sRes1 = base + C::OFF_RES1, sRes2 = base + C::OFF_RES2, sRing = base + C::OFF_RING;
  const uint32_t sDS = base + C::OFF_DS, sP = base + C::OFF_P, bar = base + C::OFF_BAR;
  const uint32_t res_full = bar, res_empty = bar + 8, s_full = bar + 16, p_full = bar + 24, acc_full = bar + 32;
  auto full = [&](int i) { return bar + 40u + 8u * i; };
  auto empty = [&](int i) { return bar + 40u + 8u * (NSLOT + i); };
  const uint32_t tmem_slot = bar + 40u + 8u * 2 * NSLOT;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_BAR + 40 + 8 * 2 * NSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (Lstr + BS - 1) / BS;
  const int rtiles = (Lres + BR - 1) / BR;
  const int n_items = rtiles * heads * samples;            // work item w -> (resident tile fastest, head, sample)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_res1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_res2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_str1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_str2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_out1) : "memory");
    if (DKDV) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_out2) : "memory");
    mbar_init(res_full, 1);
    mbar_init(res_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, SM_WARPS);
    mbar_init(acc_full, 1);
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0, it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int r0 = (w % rtiles) * BR, h = (w / rtiles) % heads, b = w / (rtiles * heads);
        mbar_wait(res_empty, ((uint32_t)it & 1u) ^ 1u);      // previous item's first-stage MMAs have consumed the tiles
        mbar_expect_tx(res_full, 2 * RES_BYTES);
        tma_load_3d(sRes1, &tm_res1, res_full, h * HD, r0, b);
        tma_load_3d(sRes2, &tm_res2, res_full, h * HD, r0, b);
        for (int j = 0; j < nt; ++j, ++g) {
          const int slot = g % NSLOT;
          mbar_wait(empty(slot), ((uint32_t)(g / NSLOT) & 1u) ^ 1u);
          mbar_expect_tx(full(slot), SLOT_BYTES);
          tma_load_3d(sRing + slot * SLOT_BYTES, &tm_str1, full(slot), h * HD, j * BS, b);
          tma_load_3d(sRing + slot * SLOT_BYTES + STR_BYTES, &tm_str2, full(slot), h * HD, j * BS, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if constexpr (CONV) {
      const uint32_t leader = elect_one();
      const uint32_t id2 = idesc(HD, 1);
      const int w16_last = ((Lstr - (nt - 1) * BS) + 15) & ~15;
      const uint32_t id1_full = idesc(BS, 0), id1_last = idesc(w16_last, 0);
      int slot = 0;
      uint32_t ph = 0, g = 0, it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        mbar_wait(res_full, it & 1u);
        tc_fence_after();
        const uint64_t r1 = desc128(sRes1), r2 = desc128(sRes2);
        for (int j = 0; j < nt; ++j, ++g) {
          const bool last = j == nt - 1;
          mbar_wait(full(slot), ph);
          tc_fence_after();
          const uint32_t y1 = sRing + slot * SLOT_BYTES, y2 = y1 + STR_BYTES;
          const uint32_t id1 = last ? id1_last : id1_full;
          const uint64_t d1 = desc128(y1), d2 = desc128(y2);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) tc_mma_p(leader, tmem + T1_COL, r1 + (uint64_t)(2 * k), d1 + (uint64_t)(2 * k), id1, k != 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) tc_mma_p(leader, tmem + T2_COL, r2 + (uint64_t)(2 * k), d2 + (uint64_t)(2 * k), id1, k != 0);
          if (last) tc_commit_p(leader, res_empty);
          tc_commit_p(leader, s_full);
          mbar_wait(p_full, g & 1u);
          tc_fence_after();
          const int ksteps = last ? w16_last / 16 : BS / 16;
#pragma unroll
          for (int k = 0; k < BS / 16; ++k) {
            if (k < ksteps) {
              if constexpr (DKDV) {
                tc_mma_p(leader, tmem + ACC1_COL, desc128(sP + (uint32_t)(k * 32)), desc128(y2 + (uint32_t)(k * 2048)), id2, (uint32_t)(j | k));
                tc_mma_p(leader, tmem + ACC2_COL, desc128(sDS + (uint32_t)(k * 32)), desc128(y1 + (uint32_t)(k * 2048)), id2, (uint32_t)(j | k));
              } else {
                tc_mma_p(leader, tmem + ACC1_COL, desc128(sDS + (uint32_t)(k * 32)), desc128(y1 + (uint32_t)(k * 2048)), id2, (uint32_t)(j | k));
              }
            }
          }
          tc_commit_p(leader, empty(slot));
          if (last) tc_commit_p(leader, acc_full);
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
    } else if (lane == 0) {
      int g = 0, it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        mbar_wait(res_full, (uint32_t)it & 1u);
        tc_fence_after();
        const uint64_t r1 = desc128(sRes1), r2 = desc128(sRes2);
        for (int j = 0; j < nt; ++j, ++g) {
          const int slot = g % NSLOT;
          int wv = Lstr - j * BS;
          wv = wv < BS ? wv : BS;
          const int w16 = (wv + 15) & ~15;                    // streamed rows in this tile, rounded to the MMA K/N step
          mbar_wait(full(slot), (uint32_t)(g / NSLOT) & 1u);
          tc_fence_after();
          const uint32_t y1 = sRing + slot * SLOT_BYTES, y2 = y1 + STR_BYTES;
          const uint32_t id1 = idesc(w16, 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) tc_mma(tmem + T1_COL, r1 + (uint64_t)(2 * k), desc128(y1) + (uint64_t)(2 * k), id1, k != 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) tc_mma(tmem + T2_COL, r2 + (uint64_t)(2 * k), desc128(y2) + (uint64_t)(2 * k), id1, k != 0);
          if (j == nt - 1) tc_commit(res_empty);
          tc_commit(s_full);
          mbar_wait(p_full, (uint32_t)g & 1u);                // P / dS in smem, T1 / T2 consumed
          tc_fence_after();
          const uint32_t id2 = idesc(HD, 1);
          const int ksteps = w16 / 16;
          for (int k = 0; k < ksteps; ++k) {                  // A: +32 B per 16 streamed rows; B (MN-major): +2 KiB
            if constexpr (DKDV) {
              tc_mma(tmem + ACC1_COL, desc128(sP + (uint32_t)(k * 32)), desc128(y2 + (uint32_t)(k * 2048)), id2, (j | k) != 0);
              tc_mma(tmem + ACC2_COL, desc128(sDS + (uint32_t)(k * 32)), desc128(y1 + (uint32_t)(k * 2048)), id2, (j | k) != 0);
            } else {
              tc_mma(tmem + ACC1_COL, desc128(sDS + (uint32_t)(k * 32)), desc128(y1 + (uint32_t)(k * 2048)), id2, (j | k) != 0);
            }
          }
          tc_commit(empty(slot));
          if (j == nt - 1) tc_commit(acc_full);
        }
      }
    }
  } else {
    // ===================== recompute / epilogue (8 warps, two threads per TMEM lane) =====================
    const int sw = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are visible to this warp
    const int hh = sw >> 2;                       // which 32-column half of the 64 streamed rows / output columns
    const int r = quarter * 32 + lane;
    const int tid = sw * 32 + lane;               // 0..255
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float2* cstat = reinterpret_cast<float2*>(smem_gen + C::OFF_STAT);
    int g = 0, it = 0;
    uint32_t dseed = 0;
    if constexpr (DROP) dseed = drop_site_seed(rng_state, drop_site);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int r0 = (w % rtiles) * BR, h = (w / rtiles) % heads, b = w / (rtiles * heads);
      const float2* st = stats + ((int64_t)b * heads + h) * Lq;
      // dropout-mask coordinates: a = (sample, head, query) row id (x C1), b = key index (x C2)
      const uint32_t row_base = (uint32_t)((b * heads + h) * Lq);
      const uint32_t fixed = DKDV ? dseed ^ (uint32_t)(r0 + r) * kDropC2 : dseed ^ (row_base + (uint32_t)(r0 + r)) * kDropC1;
      float lse_r = 0.f, d_r = 0.f;
      if constexpr (!DKDV) {
        if (r0 + r < Lq) { const float2 s = __ldg(st + r0 + r); lse_r = s.x; d_r = s.y; }
      } else {
        if (tid < BS) cstat[(g & 1) * BS + tid] = tid < Lstr ? __ldg(st + tid) : make_float2(0.f, 0.f);
      }
      for (int j = 0; j < nt; ++j, ++g) {
        float2 nxt = make_float2(0.f, 0.f);
        if constexpr (DKDV) {                                  // prefetch the next tile's column statistics
          const int q = (j + 1) * BS + tid;
          if (tid < BS && j + 1 < nt && q < Lstr) nxt = __ldg(st + q);
        }
        int wv = Lstr - j * BS;
        wv = wv < BS ? wv : BS;
        const int w16 = (wv + 15) & ~15;
        mbar_wait(s_full, (uint32_t)g & 1u);
        tc_fence_after();
        if constexpr (DKDV) asm volatile("bar.sync 1, 256;" ::: "memory");     // cstat[g & 1] is visible
        const uint32_t prow = sP + (uint32_t)(r * 128), dsrow = sDS + (uint32_t)(r * 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = hh * 32 + half * 16;
          if (c0 < w16) {                                      // warp-uniform
            uint32_t t1[16], t2[16];
            tc_ld16(lane_addr + T1_COL + c0, t1);
            tc_ld16(lane_addr + T2_COL + c0, t2);
            tc_wait_ld();
            const uint32_t var0 = DKDV ? row_base + (uint32_t)(j * BS + c0) : (uint32_t)(j * BS + c0);
            recompute16<DKDV, DROP>(t1, t2, scale_log2, cstat + (g & 1) * BS + c0, lse_r, d_r, prow, dsrow, c0 >> 3, r, fixed, var0,
                                    DKDV ? kDropC1 : kDropC2, drop_thr, drop_rk);
          }
        }
        if constexpr (DKDV) {
          if (tid < BS) cstat[((g + 1) & 1) * BS + tid] = nxt;  // readers of that buffer finished before this tile's bar.sync
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: accumulators -> bf16 -> swizzled staging (the dS / P tiles: every MMA has retired) -> TMA store
      mbar_wait(acc_full, (uint32_t)it & 1u);
      tc_fence_after();
      if constexpr (DKDV) {
        stage32(lane_addr + ACC1_COL + hh * 32, 1.0f, sP + (uint32_t)(r * 128), hh * 4, r);        // dV
        stage32(lane_addr + ACC2_COL + hh * 32, scale, sDS + (uint32_t)(r * 128), hh * 4, r);      // dK
      } else {
        stage32(lane_addr + ACC1_COL + hh * 32, scale, sDS + (uint32_t)(r * 128), hh * 4, r);      // dQ
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        if constexpr (DKDV) {
          tma_store_3d(&tm_out1, sP, h * HD, r0, b);
          tma_store_3d(&tm_out2, sDS, h * HD, r0, b);
        } else {
          tma_store_3d(&tm_out1, sDS, h * HD, r0, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// (lse, D = rowsum(dO o O)) per (sample, head, query): 8 lanes per (row, head), 16 bytes of O and dO each
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ O, int64_t ldo, int64_t obs,
                                                             const __nv_bfloat16* __restrict__ dO, int64_t ldg, int64_t gbs,
                                                             const float* __restrict__ lse, float2* __restrict__ stats,
                                                             int samples, int heads, int Lq) {
  const int64_t unit = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;    // (b, q, h) with h fastest
  const int sub = threadIdx.x & 7;
  const int64_t total = (int64_t)samples * Lq * heads;
  float acc = 0.f;
  int h = 0, q = 0, b = 0;
  if (unit < total) {
    h = (int)(unit % heads);
    q = (int)((unit / heads) % Lq);
    b = (int)(unit / ((int64_t)heads * Lq));
    const uint4 o4 = __ldg(reinterpret_cast<const uint4*>(O + b * obs + (int64_t)q * ldo + h * HD + sub * 8));
    const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(dO + b * gbs + (int64_t)q * ldg + h * HD + sub * 8));
    const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o4);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __bfloat1622float2(o2[i]), c = __bfloat1622float2(g2[i]);
      acc = fmaf(a.x, c.x, acc);
      acc = fmaf(a.y, c.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (unit < total && sub == 0) {
    const int64_t i = ((int64_t)b * heads + h) * Lq + q;
    stats[i] = make_float2(__ldg(lse + i), acc);
  }
}

template <bool DKDV, bool DROP, bool CONV>
static int configure() {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_tc_kernel<DKDV, DROP, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg<DKDV>::SMEM);
    if (e != cudaSuccess) { set_error("attention_bwd_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    done = true;
  }
  return TCD_OK;
}

}  // namespace fab

int attention_bf16_tc(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs, const void* V,
                      int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                      float scale, float* lse, float dropout_p, const void* rng_state, uint32_t site, cudaStream_t st);

}  // namespace tcd

using namespace tcd;

extern "C" int64_t tcd_attention_train_workspace_floats(int samples, int heads, int Lq) {
  return 2 * (int64_t)samples * heads * Lq;      // (lse, D) pairs of the backward pass
}

extern "C" int tcd_attention_train_forward(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs,
                                           const void* V, int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs,
                                           float* lse, int samples, int heads, int Lq, int Lk, float scale, float dropout_p,
                                           const void* rng_state, uint32_t site, void* stream) {
  TCD_REQUIRE(samples >= 0 && heads > 0 && Lq >= 0 && Lk > 0, "tcd_attention_train_forward: bad shape");
  if (samples == 0 || Lq == 0) return TCD_OK;
  TCD_REQUIRE(Q && K && V && O && lse, "tcd_attention_train_forward: null pointer");
  return attention_bf16_tc(Q, ldq, qbs, K, ldk, kbs, V, ldv, vbs, O, ldo, obs, samples, heads, Lq, Lk, scale, lse, dropout_p,
                           rng_state, site, as_stream(stream));
}

extern "C" int tcd_attention_train_backward(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs,
                                            const void* V, int64_t ldv, int64_t vbs, const void* O, int64_t ldo, int64_t obs,
                                            const void* dO, int64_t ldg, int64_t gbs, const float* lse, void* dQ,
                                            int64_t lddq, int64_t dqbs, void* dK, int64_t lddk, int64_t dkbs, void* dV,
                                            int64_t lddv, int64_t dvbs, float* stats_ws, int samples, int heads, int Lq,
                                            int Lk, float scale, float dropout_p, const void* rng_state, uint32_t site,
                                            void* stream) {
  TCD_REQUIRE(samples >= 0 && heads > 0 && Lq >= 0 && Lk > 0, "tcd_attention_train_backward: bad shape");
  if (samples == 0 || Lq == 0) return TCD_OK;
  TCD_REQUIRE(Q && K && V && O && dO && lse && dQ && dK && dV && stats_ws, "tcd_attention_train_backward: null pointer");
  const int64_t all = ldq | qbs | ldk | kbs | ldv | vbs | ldo | obs | ldg | gbs | lddq | dqbs | lddk | dkbs | lddv | dvbs;
  TCD_REQUIRE(all % 8 == 0, "tcd_attention_train_backward: pitches and batch strides must be multiples of 8 elements");
  const uintptr_t ptrs = (uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O | (uintptr_t)dO | (uintptr_t)dQ |
                         (uintptr_t)dK | (uintptr_t)dV;
  TCD_REQUIRE(ptrs % 16 == 0 && (uintptr_t)stats_ws % 8 == 0, "tcd_attention_train_backward: pointer alignment");
  TCD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || rng_state), "tcd_attention_train_backward: bad dropout arguments");
  const bool drop = dropout_p > 0.f;
  const uint32_t thr = drop ? drop_threshold(dropout_p) : 0u;
  const float rk = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;
  const uint64_t* rs = (const uint64_t*)rng_state;
  cudaStream_t st = as_stream(stream);
  float2* stats = reinterpret_cast<float2*>(stats_ws);
  {
    const int64_t threads = (int64_t)samples * Lq * heads * 8;
    fab::attn_bwd_delta_kernel<<<(unsigned)ceil_div(threads, (int64_t)256), 256, 0, st>>>(
        (const __nv_bfloat16*)O, ldo, obs, (const __nv_bfloat16*)dO, ldg, gbs, lse, stats, samples, heads, Lq);
    int rc = check_launch("attn_bwd_delta");
    if (rc) return rc;
  }
  const int64_t cols = (int64_t)heads * fab::HD;
  const int resident = 2 * num_sms();
  constexpr bool conv = kTrainConvDefault;             // converged issue loops (r01 A/B)
  int rc;
  {  // dK, dV
    CUtensorMap tk, tv, tq, tg, tdv, tdk;
    if ((rc = make_tmap_3d_bf16(&tk, K, cols, Lk, samples, ldk, kbs, fab::BR))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv, V, cols, Lk, samples, ldv, vbs, fab::BR))) return rc;
    if ((rc = make_tmap_3d_bf16(&tq, Q, cols, Lq, samples, ldq, qbs, fab::BS))) return rc;
    if ((rc = make_tmap_3d_bf16(&tg, dO, cols, Lq, samples, ldg, gbs, fab::BS))) return rc;
    if ((rc = make_tmap_3d_bf16(&tdv, dV, cols, Lk, samples, lddv, dvbs, fab::BR))) return rc;
    if ((rc = make_tmap_3d_bf16(&tdk, dK, cols, Lk, samples, lddk, dkbs, fab::BR))) return rc;
    const int64_t items = (int64_t)ceil_div(Lk, fab::BR) * heads * samples;
    const int grid = (int)(items < resident ? items : resident);
    if (drop) {
      if (conv) {
        if ((rc = fab::configure<true, true, true>())) return rc;
        fab::attention_bwd_tc_kernel<true, true, true><<<grid, fab::THREADS, fab::Cfg<true>::SMEM, st>>>(
          tk, tv, tq, tg, tdv, tdk, stats, Lk, Lq, Lq, heads, samples, scale * 1.4426950408889634f, scale, thr, rk, rs, site);
      } else {
        if ((rc = fab::configure<true, true, false>())) return rc;
        fab::attention_bwd_tc_kernel<true, true, false><<<grid, fab::THREADS, fab::Cfg<true>::SMEM, st>>>(
          tk, tv, tq, tg, tdv, tdk, stats, Lk, Lq, Lq, heads, samples, scale * 1.4426950408889634f, scale, thr, rk, rs, site);
      }
    } else {
      if (conv) {
        if ((rc = fab::configure<true, false, true>())) return rc;
        fab::attention_bwd_tc_kernel<true, false, true><<<grid, fab::THREADS, fab::Cfg<true>::SMEM, st>>>(
          tk, tv, tq, tg, tdv, tdk, stats, Lk, Lq, Lq, heads, samples, scale * 1.4426950408889634f, scale, 0u, 1.0f, nullptr, 0u);
      } else {
        if ((rc = fab::configure<true, false, false>())) return rc;
        fab::attention_bwd_tc_kernel<true, false, false><<<grid, fab::THREADS, fab::Cfg<true>::SMEM, st>>>(
          tk, tv, tq, tg, tdv, tdk, stats, Lk, Lq, Lq, heads, samples, scale * 1.4426950408889634f, scale, 0u, 1.0f, nullptr, 0u);
      }
    }
    if ((rc = check_launch("attention_bwd_tc<dkdv>"))) return rc;
  }
  {  // dQ
    CUtensorMap tq, tg, tk, tv, tdq;
    if ((rc = make_tmap_3d_bf16(&tq, Q, cols, Lq, samples, ldq, qbs, fab::BR))) return rc;
    if ((rc = make_tmap_3d_bf16(&tg, dO, cols, Lq, samples, ldg, gbs, fab::BR))) return rc;
    if ((rc = make_tmap_3d_bf16(&tk, K, cols, Lk, samples, ldk, kbs, fab::BS))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv, V, cols, Lk, samples, ldv, vbs, fab::BS))) return rc;
    if ((rc = make_tmap_3d_bf16(&tdq, dQ, cols, Lq, samples, lddq, dqbs, fab::BR))) return rc;
    const int64_t items = (int64_t)ceil_div(Lq, fab::BR) * heads * samples;
    const int grid = (int)(items < resident ? items : resident);
    if (drop) {
      if (conv) {
        if ((rc = fab::configure<false, true, true>())) return rc;
        fab::attention_bwd_tc_kernel<false, true, true><<<grid, fab::THREADS, fab::Cfg<false>::SMEM, st>>>(
          tq, tg, tk, tv, tdq, tdq, stats, Lq, Lk, Lq, heads, samples, scale * 1.4426950408889634f, scale, thr, rk, rs, site);
      } else {
        if ((rc = fab::configure<false, true, false>())) return rc;
        fab::attention_bwd_tc_kernel<false, true, false><<<grid, fab::THREADS, fab::Cfg<false>::SMEM, st>>>(
          tq, tg, tk, tv, tdq, tdq, stats, Lq, Lk, Lq, heads, samples, scale * 1.4426950408889634f, scale, thr, rk, rs, site);
      }
    } else {
      if (conv) {
        if ((rc = fab::configure<false, false, true>())) return rc;
        fab::attention_bwd_tc_kernel<false, false, true><<<grid, fab::THREADS, fab::Cfg<false>::SMEM, st>>>(
          tq, tg, tk, tv, tdq, tdq, stats, Lq, Lk, Lq, heads, samples, scale * 1.4426950408889634f, scale, 0u, 1.0f, nullptr, 0u);
      } else {
        if ((rc = fab::configure<false, false, false>())) return rc;
        fab::attention_bwd_tc_kernel<false, false, false><<<grid, fab::THREADS, fab::Cfg<false>::SMEM, st>>>(
          tq, tg, tk, tv, tdq, tdq, stats, Lq, Lk, Lq, heads, samples, scale * 1.4426950408889634f, scale, 0u, 1.0f, nullptr, 0u);
      }
    }
    if ((rc = check_launch("attention_bwd_tc<dq>"))) return rc;
  }
  return TCD_OK;
}
