// bf16 flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 64, no mask:
//   O = softmax(scale * Q K^T) V   per (sample, head);  replaces SBI_MSA's core (model/model.py:97-102)
//   and the nn.MultiheadAttention core of the music encoder in bf16 mode.
//
// Persistent CTAs (2 per SM: ~97 KiB smem, 256 TMEM columns and <= 96 registers each) loop over work items = one
// 128-query tile of one (sample, head).  Keys are consumed 64 at a time:
//   warp 0      TMA producer: Q tile per item, then K(0) V(0) K(1) V(1) ... as 64-key x 64 bf16 boxes (3-D tensor
//               maps (col, row, sample): rows past Lq/Lk are zero-filled, never read from the next sample) into a
//               6-slot shared-memory ring that keeps prefetching across item boundaries
//   warp 1      MMA issuer (converged warp, one elected lane issues).  S(t) = Q K(t)^T : tcgen05.mma M=128 N=kw K=16 x4 into TMEM buffer t%2 —
//               issued ONE TILE AHEAD of the softmax, so the softmax warps never wait for the tensor core;
//               O += P(t) V(t) : M=128 N=64 K=16 x kw/16, V as MN-major operand, accumulating in TMEM.
//               TMEM: S0 [0,64) | S1 [64,128) | O [128,192)
//   warps 2..9  softmax: two threads per query row (TMEM lane), 32 keys each: ONE tcgen05.ld of the thread's
//               scores, row max exchanged with the partner through (double-buffered) smem, p = exp2(s*c - m),
//               P(t) as bf16 into the 128B-swizzled P buffer t%2 (A operand of P V).  The output row lives in
//               TMEM and is rescaled (tcgen05.ld / multiply / tcgen05.st) only when the row's reference max has to
//               move, which is deferred until the running max grows by more than 2^8 — exact, because P, l and O
//               share the reference.  Final O / l goes through smem and one TMA bulk store (rows past Lq clipped).
// Keys past Lk in the last tile are masked to -inf; the last tile's MMA width kw is only rounded up to 16 keys.
// r01 history (self-attention L=750, TFLOP/s): thread-per-row 306 -> two threads per row 414 -> S(t+1) issued
// before P V(t) 435 -> persistent CTAs (Lk=152: 211 -> 268) -> S double-buffered, O in TMEM 464 -> converged MMA
// issue loop 500 (Lk=152: 298).
#include "tc_attn_common.cuh"
#include "dropout.cuh"
#include "tuning.cuh"

namespace tcd {

namespace fa {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = 128 * 128;              // 128 rows x 64 bf16
constexpr int KV_BYTES = BKV * 128;             // 64 rows x 64 bf16
constexpr int P_BYTES = 128 * 128;              // 128 rows x 64 keys bf16
constexpr int NSLOT = 6;
constexpr int SM_WARPS = 8;
constexpr int THREADS = (2 + SM_WARPS) * 32;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128;           // S buffers at S_COL + 64*b
// smem: Q | ring[6] | P[2] (P[0] doubles as the output staging tile) | max/sum exchange [2][128][2] | barriers
constexpr int OFF_Q = 0, OFF_RING = Q_BYTES, OFF_P = OFF_RING + NSLOT * KV_BYTES, OFF_X = OFF_P + 2 * P_BYTES;
constexpr int OFF_BAR = OFF_X + 2 * 128 * 2 * 4;
constexpr size_t SMEM = 1024 + OFF_BAR + 256;

// softmax of one thread's 32 scores.  MASKED only for the last (partial) tile of a row of keys.
template <bool MASKED>
__device__ __forceinline__ float row_max32(const uint32_t (&raw)[32], int valid) {
  float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    a0 = fmaxf(a0, (!MASKED || j < valid) ? __uint_as_float(raw[j]) : -INFINITY);
    a1 = fmaxf(a1, (!MASKED || j + 1 < valid) ? __uint_as_float(raw[j + 1]) : -INFINITY);
    a2 = fmaxf(a2, (!MASKED || j + 2 < valid) ? __uint_as_float(raw[j + 2]) : -INFINITY);
    a3 = fmaxf(a3, (!MASKED || j + 3 < valid) ? __uint_as_float(raw[j + 3]) : -INFINITY);
  }
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}
// DROP: the stored probabilities (the P V operand) are multiplied by the attention-dropout mask / (1-p)
// (model/model.py:98, nn.MultiheadAttention dropout); the row sum keeps the undropped softmax normalisation.
template <bool MASKED, bool DROP>
__device__ __forceinline__ float exp_store32(const uint32_t (&raw)[32], int valid, float scale_log2, float mt, uint32_t rowb,
                                             int chunk0, int r, uint32_t rowseed, uint32_t key0, uint32_t thr, float rk) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float p[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sc = (!MASKED || 8 * j + e < valid) ? __uint_as_float(raw[8 * j + e]) : -INFINITY;
      p[e] = ex2(fmaf(sc, scale_log2, -mt));               // -inf -> 0
    }
    s0 += p[0] + p[4]; s1 += p[1] + p[5]; s2 += p[2] + p[6]; s3 += p[3] + p[7];
    if constexpr (DROP) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        p[e] = drop_mix(rowseed ^ (key0 + (uint32_t)(8 * j + e)) * kDropC2) >= thr ? p[e] * rk : 0.f;
    }
    sts128(rowb + (uint32_t)((((chunk0 + j) ^ r) & 7) << 4), pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]),
           pack2(p[6], p[7]));
  }
  return (s0 + s1) + (s2 + s3);
}

// Design points kept from the r01 A/B runs (profiles/r01_issue_loops.md; the losing variants are gone from the tree):
//   * the MMA issue loop runs converged (all lanes, the elected lane issues inside the asm; see the MMA role);
//   * the row-max / row-sum exchange synchronises only the two warps that share a row (named barriers 2..5, 64
//     threads) instead of all eight softmax warps;
//   * no wait on o_full before P(t) overwrites the buffer P V(t-2) read: s_full of S(t), already observed, was
//     committed after P V(t-2) by the same thread, and tcgen05.commit covers every earlier MMA of that thread.
template <bool DROP>
__global__ void __launch_bounds__(THREADS, 2) attention_tc_kernel(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, int Lq, int Lk, int heads,
    int samples, float scale_log2, float* __restrict__ lse, uint32_t drop_thr, float drop_rk,
    const uint64_t* __restrict__ rng_state, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base + OFF_Q, sRing = base + OFF_RING, sP = base + OFF_P, bar = base + OFF_BAR;
  // barriers (8 bytes each)
  const uint32_t q_full = bar, q_empty = bar + 8;
  auto s_full = [&](int b) { return bar + 16u + 8u * b; };
  auto p_full = [&](int b) { return bar + 32u + 8u * b; };
  auto o_full = [&](int b) { return bar + 48u + 8u * b; };
  auto full = [&](int i) { return bar + 64u + 8u * i; };
  auto empty = [&](int i) { return bar + 64u + 8u * (NSLOT + i); };
  const uint32_t tmem_slot = bar + 64u + 8u * 2 * NSLOT;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 64 + 8 * 2 * NSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (Lk + BKV - 1) / BKV;
  const int qtiles = (Lq + BQ - 1) / BQ;
  const int n_items = qtiles * heads * samples;          // work item w -> (q tile fastest, head, sample)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o) : "memory");
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(p_full(b), SM_WARPS); mbar_init(o_full(b), 1); }
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
      int g = 0;                                            // ring item counter across work items
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int q0 = (w % qtiles) * BQ, h = (w / qtiles) % heads, b = w / (qtiles * heads);
        mbar_wait(q_empty, ((uint32_t)it & 1u) ^ 1u);       // previous item's last S MMA has consumed Q
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_3d(sQ, &tm_q, q_full, h * HD, q0, b);
        for (int item = 0; item < 2 * nt; ++item, ++g) {    // K(0) V(0) K(1) V(1) ...
          const int slot = g % NSLOT;
          const uint32_t ph = (uint32_t)(g / NSLOT) & 1u;
          mbar_wait(empty(slot), ph ^ 1u);
          mbar_expect_tx(full(slot), KV_BYTES);
          tma_load_3d(sRing + slot * KV_BYTES, (item & 1) ? &tm_v : &tm_k, full(slot), h * HD, (item >> 1) * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      // Converged issue loop.  With the loop under `if (lane == 0)` ptxas keeps every descriptor in vector
      // registers and wraps each UTCHMMA / UTCBAR in an R2UR + ELECT "waterfall" loop: ~250 dependent single-thread
      // instructions per key tile, which (not the softmax) paced the kernel.  Here all lanes run the warp-uniform
      // loop and one elected lane issues; ring positions are carried as (slot, phase) counters instead of g / 6.
      const uint32_t leader = elect_one();
      const uint64_t qdesc = desc128(sQ);
      const int kw_last = ((Lk - (nt - 1) * BKV) + 15) & ~15;   // MMA width of an item's last key tile
      const uint32_t id_s_full = idesc(BKV, 0), id_s_last = idesc(kw_last, 0), id_pv = idesc(HD, 1);
      const int ksteps_last = kw_last / 16;
      int ks = 0, vs = 1;                                       // ring slots of the next K (even) / V (odd) tile
      uint32_t kph = 0, vph = 0, tcg = 0, it = 0;               // their phases; KV-tile and work-item counters
      auto issue_s = [&](bool last, uint32_t sbuf) {
        mbar_wait(full(ks), kph);
        tc_fence_after();
        const uint64_t kdesc = desc128(sRing + ks * KV_BYTES);
        const uint32_t id = last ? id_s_last : id_s_full;
        const uint32_t d = tmem + S_COL + 64u * sbuf;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_p(leader, d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), id, k != 0);
        tc_commit_p(leader, empty(ks));
        if (last) tc_commit_p(leader, q_empty);
        tc_commit_p(leader, s_full(sbuf));
        ks += 2;
        if (ks == NSLOT) { ks = 0; kph ^= 1u; }
      };
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1u);
        tc_fence_after();
        issue_s(nt == 1, tcg & 1u);
        for (int t = 0; t < nt; ++t, ++tcg) {
          const uint32_t b = tcg & 1u;
          if (t + 1 < nt) issue_s(t + 2 == nt, b ^ 1u);         // one tile ahead of the softmax
          mbar_wait(p_full(b), (tcg >> 1) & 1u);
          tc_fence_after();
          mbar_wait(full(vs), vph);
          tc_fence_after();
          const uint32_t vbase = sRing + vs * KV_BYTES, pbase = sP + b * P_BYTES;
          const int ksteps = (t == nt - 1) ? ksteps_last : BKV / 16;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            if (k < ksteps)
              tc_mma_p(leader, tmem + O_COL, desc128(pbase + (uint32_t)(k * 32)), desc128(vbase + (uint32_t)(k * 2048)), id_pv,
                       (uint32_t)(t | k));
          tc_commit_p(leader, empty(vs));
          tc_commit_p(leader, o_full(b));
          vs += 2;
          if (vs > NSLOT) { vs = 1; vph ^= 1u; }
        }
      }
    }
  } else {
    // ===================== softmax / output (8 warps, two threads per query row) =====================
    const int sw = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are visible to this warp
    const int hh = sw >> 2;                       // which 32-key half of the tile / 32-column half of the output
    const int r = quarter * 32 + lane;
    auto pair_sync = [&]() {
      if (quarter == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
      else if (quarter == 1) asm volatile("bar.sync 3, 64;" ::: "memory");
      else if (quarter == 2) asm volatile("bar.sync 4, 64;" ::: "memory");
      else asm volatile("bar.sync 5, 64;" ::: "memory");
    };
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem_gen + OFF_X);
    int tc = 0;                                            // KV-tile counter across work items
    uint32_t dseed = 0;
    if constexpr (DROP) dseed = drop_site_seed(rng_state, drop_site);
    // final O (TMEM) / l -> bf16 -> swizzled staging tile (P buffer 0; every P V has retired) -> TMA store.
    //   tcl = the item's last key tile; tnext = the next tile this CTA will process (its max-exchange buffer is idle)
    auto finish = [&](int w_e, float m_e, float l_e, int tcl, int tnext) {
      const int q0 = (w_e % qtiles) * BQ, h = (w_e / qtiles) % heads, b = w_e / (qtiles * heads);
      mbar_wait(o_full(tcl & 1), (uint32_t)(tcl >> 1) & 1u);
      tc_fence_after();
      uint32_t ov[32];
      tc_ld32(lane_addr + O_COL + hh * 32, ov);
      tc_wait_ld();
      tc_fence_before();
      float* xs = xch + (tnext & 1) * 256;
      xs[r * 2 + hh] = l_e;
      pair_sync();
      const float lsum = xs[r * 2] + xs[r * 2 + 1];
      const float inv = 1.0f / lsum;
      // log2-domain log-sum-exp of the scaled scores (training: the backward pass recomputes P = exp2(s*c - lse))
      if (lse != nullptr && hh == 0 && q0 + r < Lq) lse[((int64_t)b * heads + h) * Lq + q0 + r] = m_e + log2f(lsum);
      const uint32_t rowo = sP + (uint32_t)(r * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(rowo + (uint32_t)((((hh * 4 + j) ^ r) & 7) << 4),
               pack2(__uint_as_float(ov[8 * j]) * inv, __uint_as_float(ov[8 * j + 1]) * inv),
               pack2(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv),
               pack2(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv),
               pack2(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight softmax warps only
      if (sw == 0 && lane == 0) {
        tma_store_3d(&tm_o, sP, h * HD, q0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile is P buffer 0 of the next item
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int q0 = (w % qtiles) * BQ, h = (w / qtiles) % heads, b = w / (qtiles * heads);
      const uint32_t rowseed = dseed ^ (uint32_t)((b * heads + h) * Lq + q0 + r) * kDropC1;
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < nt; ++t, ++tc) {
        const int sb = tc & 1;
        const int valid = min(BKV, Lk - t * BKV) - hh * 32;    // valid keys among this thread's 32 (may be <= 0)
        mbar_wait(s_full(sb), (uint32_t)(tc >> 1) & 1u);
        tc_fence_after();
        uint32_t raw[32];
        float mx = -INFINITY;
        if (valid > 0) {                                       // warp-uniform
          tc_ld32(lane_addr + S_COL + 64 * sb + hh * 32, raw);
          tc_wait_ld();
          mx = valid >= 32 ? row_max32<false>(raw, 32) : row_max32<true>(raw, valid);
        }
        float* xb = xch + sb * 256;
        xb[r * 2 + hh] = mx;
        pair_sync();
        const float tile_max = fmaxf(xb[r * 2], xb[r * 2 + 1]) * scale_log2;
        // lazy reference max: move it only when the row max grew by more than 2^8 (both threads of a row agree)
        const float mt = (t == 0 || tile_max > m + 8.0f) ? tile_max : m;
        const bool moved = (t > 0) && (mt != m);
        const float corr = moved ? ex2(m - mt) : 1.0f;
        if (__any_sync(0xffffffffu, moved)) {                  // rare: rescale this warp's rows of O in TMEM
          mbar_wait(o_full((tc - 1) & 1), (uint32_t)((tc - 1) >> 1) & 1u);     // every earlier P V has retired
          tc_fence_after();
          uint32_t ov[32];
          tc_ld32(lane_addr + O_COL + hh * 32, ov);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * corr);
          tc_st32(lane_addr + O_COL + hh * 32, ov);
          tc_wait_st();
        }
        l *= corr;
        m = mt;
        // p = exp2(s*scale - m), partial row sum, bf16 P(t) into swizzled smem (row r, 16-byte chunk j at j ^ (r & 7))
        if (valid > 0) {
          const uint32_t rowb = sP + (uint32_t)(sb * P_BYTES + r * 128);
          const uint32_t key0 = (uint32_t)(t * BKV + hh * 32);
          l += valid >= 32 ? exp_store32<false, DROP>(raw, 32, scale_log2, mt, rowb, hh * 4, r, rowseed, key0, drop_thr, drop_rk)
                           : exp_store32<true, DROP>(raw, valid, scale_log2, mt, rowb, hh * 4, r, rowseed, key0, drop_thr, drop_rk);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(sb));
      }
      finish(w, m, l, tc - 1, tc);
    }  // work items
    if (sw == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace fa

// =====================================================================================================================
// Two query tiles per CTA with ping-pong softmax (round 2).
// ncu r02 of the kernel above: 9.4 issued instructions per score (the per-tile overhead — barrier waits, the row-max
// exchange between the two threads of a row, address arithmetic, fences — equals the useful work at 32 scores per thread),
// issue slots 59 % busy, MUFU 53 %, tensor pipe 25 %.  Here ONE CTA per SM owns TWO 128-query tiles of the same (sample, head):
//   * a softmax thread owns a whole query row of its tile (64 scores per key tile): row max and row sum are thread-local —
//     no shared-memory exchange, no named barrier per key tile, 5.6 instructions per score;
//   * both tiles multiply against the same K / V tile in shared memory: one TMA load feeds 256 queries;
//   * P never touches shared memory: the bf16 probabilities are written back into the tensor-memory columns of the scores
//     they came from (tcgen05.st, two keys per 32-bit column) and P V runs with its A operand in tensor memory — no
//     st.shared, no generic->async proxy fence per tile, no P buffers;
//   * the two tiles take turns entering their exp2 phase (named-barrier hand-shake between the two groups of four warps:
//     a tile may start exp2(t) only after the other tile has started its own exp2 of the same key tile, alternately).
//     Left alone the two groups drift into lockstep — both in the exp2 phase at once, then both in the latency-bound rest
//     (TMEM load, row max, P store, barrier traffic) with the MUFU idle.  Measured r02 (self-attention, 128 x 8 x 750^2):
//     free-running 278 us, exclusive exp2 windows 262 us (a warp issuing MUFU.EX2 back to back stalls 8 cycles per
//     instruction and cannot issue its own FMA work meanwhile, so a lone warp needs ~780 cycles per 64 scores and the
//     windows become the critical path), hand-shake only (ptxas keeps the register-only exp2 code outside the barrier
//     pair) 242 us = 610 TFLOP/s; windows pinned by data dependencies 244 us;
//   * S is double-buffered per tile (S_g[2] | O_g), S(t+1) is issued one key tile ahead; Q is double-buffered by work item.
// A three-tile variant with single-buffered S (all 512 TMEM columns could hold only that) measured 383 us against 278 us
// for the same self-attention launch: the S round trip lands in every warp's serial chain.
// Everything else (lazy rescale of O in TMEM, exact masking of the last key tile, dropout of the stored probabilities,
// log2-domain LSE for the training backward, TMA-stored output) is as in the kernel above.
namespace fa2 {
using namespace fa;

constexpr int NSLOT2 = 8;
constexpr int THREADS2 = (2 + 8) * 32;
constexpr int OFF2_Q = 0;                                      // Q[2 item buffers][2 tiles]
constexpr int OFF2_RING = 4 * Q_BYTES;                         // ring[8] of 64-key K / V tiles
constexpr int OFF2_OUT = OFF2_RING + NSLOT2 * KV_BYTES;        // output staging [2 tiles]
constexpr int OFF2_BAR = OFF2_OUT + 2 * Q_BYTES;
constexpr size_t SMEM2 = 1024 + OFF2_BAR + 512;
constexpr int O2_COL = 256;                                    // TMEM: S_g buffer b at 128 g + 64 b, O_g at 256 + 64 g

template <bool MASKED>
__device__ __forceinline__ float row_max64(const uint32_t (&raw)[64], int valid) {
  float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    a0 = fmaxf(a0, (!MASKED || j < valid) ? __uint_as_float(raw[j]) : -INFINITY);
    a1 = fmaxf(a1, (!MASKED || j + 1 < valid) ? __uint_as_float(raw[j + 1]) : -INFINITY);
    a2 = fmaxf(a2, (!MASKED || j + 2 < valid) ? __uint_as_float(raw[j + 2]) : -INFINITY);
    a3 = fmaxf(a3, (!MASKED || j + 3 < valid) ? __uint_as_float(raw[j + 3]) : -INFINITY);
  }
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}
// P(t) as bf16 pairs: pw[i] = keys (2i, 2i+1) of this thread's query row — the A operand layout of a tensor-memory MMA.
// The unmasked, undropped tile (11 of 12 key tiles of the sampler) runs on the packed fp32 pipe (fma.rn.f32x2 / add.rn.f32x2):
// 32 + 32 instead of 64 + ~70 FMA-pipe instructions around the 64 MUFU.EX2 — a single warp per scheduler owns the MUFU
// during its ping-pong window, so the window's length is its instruction count.
// NG: groups of 8 keys that can hold a valid key.  The cross-attention's third key tile holds 24 of 64 keys (Lk = 152): its
// masked tile runs NG = 4 — straight-line code over the first 32 keys, zeros for the rest, half the MUFU.EX2.  (Skipping every
// group past `valid` with a branch per group measured slower on the self-attention's 46-key tile: the branches break up the
// schedule of the unrolled loop.)
template <bool MASKED, bool DROP, int NG = 8>
__device__ __forceinline__ float exp_pack64(const uint32_t (&raw)[64], int valid, float scale_log2, float mt, uint32_t (&pw)[32],
                                            uint32_t rowseed, uint32_t key0, uint32_t thr, float rk) {
  if constexpr (!MASKED && !DROP) {
    const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-mt, -mt);
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[e] = __ffma2_rn(make_float2(__uint_as_float(raw[8 * j + 2 * e]), __uint_as_float(raw[8 * j + 2 * e + 1])), sc2, nm2);
        x[e].x = ex2(x[e].x);
        x[e].y = ex2(x[e].y);
      }
      a0 = __fadd2_rn(a0, x[0]); a1 = __fadd2_rn(a1, x[1]); a2 = __fadd2_rn(a2, x[2]); a3 = __fadd2_rn(a3, x[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) pw[4 * j + e] = pack2(x[e].x, x[e].y);
    }
    const float2 t = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
    return t.x + t.y;
  } else {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = NG; j < 8; ++j) { pw[4 * j] = 0u; pw[4 * j + 1] = 0u; pw[4 * j + 2] = 0u; pw[4 * j + 3] = 0u; }
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float sc = (!MASKED || 8 * j + e < valid) ? __uint_as_float(raw[8 * j + e]) : -INFINITY;
        p[e] = ex2(fmaf(sc, scale_log2, -mt));               // -inf -> 0
      }
      s0 += p[0] + p[4]; s1 += p[1] + p[5]; s2 += p[2] + p[6]; s3 += p[3] + p[7];
      if constexpr (DROP) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          p[e] = drop_mix(rowseed ^ (key0 + (uint32_t)(8 * j + e)) * kDropC2) >= thr ? p[e] * rk : 0.f;
      }
      pw[4 * j] = pack2(p[0], p[1]); pw[4 * j + 1] = pack2(p[2], p[3]); pw[4 * j + 2] = pack2(p[4], p[5]); pw[4 * j + 3] = pack2(p[6], p[7]);
    }
    return (s0 + s1) + (s2 + s3);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(THREADS2, 1) attention_tc2q_kernel(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, int Lq, int Lk, int heads,
    int samples, float scale_log2, float* __restrict__ lse, uint32_t drop_thr, float drop_rk,
    const uint64_t* __restrict__ rng_state, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base + OFF2_Q, sRing = base + OFF2_RING, sOut = base + OFF2_OUT, bar = base + OFF2_BAR;
  auto q_full = [&](int i) { return bar + 8u * i; };
  auto q_empty = [&](int i) { return bar + 16u + 8u * i; };
  auto s_full = [&](int g, int b) { return bar + 32u + 8u * (2 * g + b); };
  auto p_full = [&](int g, int b) { return bar + 64u + 8u * (2 * g + b); };
  auto o_full = [&](int g, int b) { return bar + 96u + 8u * (2 * g + b); };
  auto full = [&](int i) { return bar + 128u + 8u * i; };
  auto empty = [&](int i) { return bar + 128u + 8u * (NSLOT2 + i); };
  const uint32_t tmem_slot = bar + 128u + 8u * 2 * NSLOT2;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF2_BAR + 128 + 8 * 2 * NSLOT2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (Lk + BKV - 1) / BKV;
  const int qtiles = (Lq + BQ - 1) / BQ;
  const int qpairs = (qtiles + 1) / 2;
  const int n_items = qpairs * heads * samples;          // work item w -> (q-tile pair fastest, head, sample)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o) : "memory");
    for (int i = 0; i < 2; ++i) { mbar_init(q_full(i), 1); mbar_init(q_empty(i), 1); }
    for (int g = 0; g < 2; ++g)
      for (int b = 0; b < 2; ++b) { mbar_init(s_full(g, b), 1); mbar_init(p_full(g, b), 4); mbar_init(o_full(g, b), 1); }
    for (int i = 0; i < NSLOT2; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_sync();                // the prologue overlaps the previous kernel's tail; Q / K / V / O / lse only from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int gi = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int q0 = (w % qpairs) * 2 * BQ, h = (w / qpairs) % heads, b = w / (qpairs * heads);
        const int qb = it & 1;
        mbar_wait(q_empty(qb), (((uint32_t)it >> 1) & 1u) ^ 1u);   // the last S MMAs of item it-2 have consumed this buffer
        mbar_expect_tx(q_full(qb), 2 * Q_BYTES);
        tma_load_3d(sQ + qb * 2 * Q_BYTES, &tm_q, q_full(qb), h * HD, q0, b);
        tma_load_3d(sQ + qb * 2 * Q_BYTES + Q_BYTES, &tm_q, q_full(qb), h * HD, q0 + BQ, b);   // rows past Lq are zero-filled
        for (int item = 0; item < 2 * nt; ++item, ++gi) {            // K(0) V(0) K(1) V(1) ...
          const int slot = gi % NSLOT2;
          const uint32_t ph = (uint32_t)(gi / NSLOT2) & 1u;
          mbar_wait(empty(slot), ph ^ 1u);
          mbar_expect_tx(full(slot), KV_BYTES);
          tma_load_3d(sRing + slot * KV_BYTES, (item & 1) ? &tm_v : &tm_k, full(slot), h * HD, (item >> 1) * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged loop, one elected lane issues) =====================
    const uint32_t leader = elect_one();
    uint64_t qdesc0 = 0, qdesc1 = 0;
    uint32_t q_empty_bar = 0;
    const int kw_last = ((Lk - (nt - 1) * BKV) + 15) & ~15;
    const uint32_t id_s_full = idesc(BKV, 0), id_s_last = idesc(kw_last, 0), id_pv = idesc(HD, 1);
    const int ksteps_last = kw_last / 16;
    int ks = 0, vs = 1;                                        // ring slots of the next K (even) / V (odd) tile
    uint32_t kph = 0, vph = 0, tcg = 0, it = 0;
    auto issue_s = [&](bool last, uint32_t sbuf) {
      mbar_wait(full(ks), kph);
      tc_fence_after();
      const uint64_t kdesc = desc128(sRing + ks * KV_BYTES);
      const uint32_t id = last ? id_s_last : id_s_full;
      const uint32_t d0 = tmem + 64u * sbuf, d1 = d0 + 128u;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) tc_mma_p(leader, d0, qdesc0 + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), id, k != 0);
      tc_commit_p(leader, s_full(0, sbuf));
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) tc_mma_p(leader, d1, qdesc1 + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), id, k != 0);
      tc_commit_p(leader, empty(ks));
      if (last) tc_commit_p(leader, q_empty_bar);
      tc_commit_p(leader, s_full(1, sbuf));
      ks += 2;
      if (ks == NSLOT2) { ks = 0; kph ^= 1u; }
    };
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const uint32_t qb = it & 1u;
      qdesc0 = desc128(sQ + qb * 2 * Q_BYTES);
      qdesc1 = desc128(sQ + qb * 2 * Q_BYTES + Q_BYTES);
      q_empty_bar = q_empty((int)qb);
      mbar_wait(q_full((int)qb), (it >> 1) & 1u);
      tc_fence_after();
      issue_s(nt == 1, tcg & 1u);
      for (int t = 0; t < nt; ++t, ++tcg) {
        const uint32_t b = tcg & 1u;
        if (t + 1 < nt) issue_s(t + 2 == nt, b ^ 1u);         // one key tile ahead of the softmax (behind P V(t-1), which read that buffer)
        const int ksteps = (t == nt - 1) ? ksteps_last : BKV / 16;
        mbar_wait(full(vs), vph);
        const uint32_t vbase = sRing + vs * KV_BYTES;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(p_full(g, b), (tcg >> 1) & 1u);           // P_g(t) in TMEM (replacing S_g(t)), O_g rescaled if needed
          tc_fence_after();
          // A = P_g(t): bf16 pairs in the first 32 columns of S buffer b of tile g (8 columns per 16 keys)
          const uint32_t ptm = tmem + 128u * g + 64u * b;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            if (k < ksteps)
              tc_mma_ts_p(leader, tmem + O2_COL + 64u * g, ptm + (uint32_t)(8 * k), desc128(vbase + (uint32_t)(k * 2048)), id_pv,
                          (uint32_t)(t | k));
          if (g == 1) tc_commit_p(leader, empty(vs));
          tc_commit_p(leader, o_full(g, b));
        }
        vs += 2;
        if (vs > NSLOT2) { vs = 1; vph ^= 1u; }
      }
    }
  } else {
    // ===================== softmax / output: warps 2-5 tile A, 6-9 tile B, one thread per query row =====================
    const int sw = warp - 2;
    const int g = sw >> 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are visible to this warp
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_col = (uint32_t)(128 * g), o_col = (uint32_t)(O2_COL + 64 * g);
    const uint32_t sOg = sOut + (uint32_t)(g * Q_BYTES);
    int tc = 0;                                            // key-tile counter across work items
    uint32_t dseed = 0;
    if constexpr (DROP) dseed = drop_site_seed(rng_state, drop_site);
    auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };   // the four warps of this tile
    // turn taking: barrier 3 + g admits tile g to its exp2 phase; the other tile's warps arrive on it when they enter theirs
    auto pp_enter = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(3 + g) : "memory"); };
    auto pp_leave = [&]() { asm volatile("bar.arrive %0, 256;" ::"r"(4 - g) : "memory"); };
    if (g == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");     // tile A goes first
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int row0 = (w % qpairs) * 2 * BQ + g * BQ, h = (w / qpairs) % heads, b = w / (qpairs * heads);
      const uint32_t rowseed = dseed ^ (uint32_t)((b * heads + h) * Lq + row0 + r) * kDropC1;
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < nt; ++t, ++tc) {
        const int sb = tc & 1;
        const int valid = min(BKV, Lk - t * BKV);          // >= 1
        mbar_wait(s_full(g, sb), (uint32_t)(tc >> 1) & 1u);
        tc_fence_after();
        uint32_t raw[64];
        {
          uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
          uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[32]);
          tc_ld32(lane_addr + s_col + 64 * sb, lo);
          tc_ld32(lane_addr + s_col + 64 * sb + 32, hi);
          tc_wait_ld();
        }
        const float mx = valid >= BKV ? row_max64<false>(raw, BKV) : row_max64<true>(raw, valid);
        const float tile_max = mx * scale_log2;
        // lazy reference max: move it only when the row max grew by more than 2^8
        const float mt = (t == 0 || tile_max > m + 8.0f) ? tile_max : m;
        const bool moved = (t > 0) && (mt != m);
        const float corr = moved ? ex2(m - mt) : 1.0f;
        if (__any_sync(0xffffffffu, moved)) {                  // rare: rescale this warp's rows of O in TMEM
          mbar_wait(o_full(g, (tc - 1) & 1), (uint32_t)((tc - 1) >> 1) & 1u);     // every earlier P V of this tile has retired
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t ov[32];
            tc_ld32(lane_addr + o_col + 32 * hf, ov);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * corr);
            tc_st32(lane_addr + o_col + 32 * hf, ov);
          }
          tc_wait_st();
        }
        l *= corr;
        m = mt;
        const uint32_t key0 = (uint32_t)(t * BKV);
        uint32_t pw[32];
        if (!DROP && valid >= BKV) {
          // the common tile (unmasked, no dropout) on the packed fp32 pipe (fma.rn.f32x2 / add.rn.f32x2): 32 + 35 instead of
          // 64 + ~70 FMA-pipe instructions around the 64 MUFU.EX2.  pp_enter / pp_leave: the turn-taking hand-shake (the
          // barrier pair orders only the two tiles' ENTRY into this phase; ptxas schedules the arithmetic around it freely)
          const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-mt, -mt);
          float2 x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            x[i] = __ffma2_rn(make_float2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1])), sc2, nm2);
          pp_enter();
          pp_leave();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            x[i].x = ex2(x[i].x);
            x[i].y = ex2(x[i].y);
          }
          float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            a0 = __fadd2_rn(a0, x[i]); a1 = __fadd2_rn(a1, x[i + 1]); a2 = __fadd2_rn(a2, x[i + 2]); a3 = __fadd2_rn(a3, x[i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) pw[i] = pack2(x[i].x, x[i].y);
          const float2 ts = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
          l += ts.x + ts.y;
        } else {
          pp_enter();
          l += valid >= BKV ? exp_pack64<false, DROP>(raw, BKV, scale_log2, mt, pw, rowseed, key0, drop_thr, drop_rk)
               : valid <= 32 ? exp_pack64<true, DROP, 4>(raw, valid, scale_log2, mt, pw, rowseed, key0, drop_thr, drop_rk)
                             : exp_pack64<true, DROP>(raw, valid, scale_log2, mt, pw, rowseed, key0, drop_thr, drop_rk);
          pp_leave();
        }
        tc_st32(lane_addr + s_col + 64 * sb, pw);              // P(t) replaces the scores this thread holds in registers
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g, sb));
      }
      // ---- final O (TMEM) / l -> bf16 -> swizzled staging tile -> TMA store
      {
        const int tcl = tc - 1;
        mbar_wait(o_full(g, tcl & 1), (uint32_t)(tcl >> 1) & 1u);
        tc_fence_after();
        const float inv = 1.0f / l;
        if (lse != nullptr && row0 + r < Lq) lse[((int64_t)b * heads + h) * Lq + row0 + r] = m + log2f(l);
        const uint32_t rowo = sOg + (uint32_t)(r * 128);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t ov[32];
          tc_ld32(lane_addr + o_col + 32 * hf, ov);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(rowo + (uint32_t)((((hf * 4 + j) ^ r) & 7) << 4),
                   pack2(__uint_as_float(ov[8 * j]) * inv, __uint_as_float(ov[8 * j + 1]) * inv),
                   pack2(__uint_as_float(ov[8 * j + 2]) * inv, __uint_as_float(ov[8 * j + 3]) * inv),
                   pack2(__uint_as_float(ov[8 * j + 4]) * inv, __uint_as_float(ov[8 * j + 5]) * inv),
                   pack2(__uint_as_float(ov[8 * j + 6]) * inv, __uint_as_float(ov[8 * j + 7]) * inv));
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        wg_sync();
        if ((sw & 3) == 0 && lane == 0) {
          if (row0 < Lq) {
            tma_store_3d(&tm_o, sOg, h * HD, row0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tile is rewritten by the next item
          }
        }
        wg_sync();
      }
    }  // work items
    if ((sw & 3) == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace fa2

template <bool DROP>
static int launch_attention_tc2q(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                                 int Lq, int Lk, int heads, int samples, float scale_log2, float* lse, uint32_t thr, float rk,
                                 const uint64_t* rng_state, uint32_t site, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fa2::attention_tc2q_kernel<DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa2::SMEM2);
    if (e != cudaSuccess) { set_error("attention_tc2q: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int64_t items = (int64_t)((ceil_div(Lq, fa::BQ) + 1) / 2) * heads * samples;
  const int grid = (int)(items < num_sms() ? items : num_sms());     // one CTA per SM (shared memory, all of TMEM)
  cudaError_t le = launch_pdl(fa2::attention_tc2q_kernel<DROP>, dim3(grid), dim3(fa2::THREADS2), fa2::SMEM2, st, items <= 2 * (int64_t)num_sms(), tq, tk, tv, to, Lq, Lk, heads,
                              samples, scale_log2, lse, thr, rk, rng_state, site);
  if (le != cudaSuccess) { set_error("attention_tc2q: launch: %s", cudaGetErrorString(le)); return TCD_ERR_CUDA; }
  return check_launch("attention_tc2q");
}

template <bool DROP>
static int launch_attention_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int grid,
                               int Lq, int Lk, int heads, int samples, float scale_log2, float* lse, uint32_t thr, float rk,
                               const uint64_t* rng_state, uint32_t site, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fa::attention_tc_kernel<DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa::SMEM);
    if (e != cudaSuccess) { set_error("attention_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  fa::attention_tc_kernel<DROP><<<grid, fa::THREADS, fa::SMEM, st>>>(tq, tk, tv, to, Lq, Lk, heads, samples, scale_log2, lse, thr,
                                                                         rk, rng_state, site);
  return check_launch("attention_tc");
}

int attention_bf16_tc(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs, const void* V,
                      int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                      float scale, float* lse, float dropout_p, const void* rng_state, uint32_t site, cudaStream_t st) {
  TCD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && qbs % 8 == 0 && kbs % 8 == 0 &&
              vbs % 8 == 0 && obs % 8 == 0, "tcd_attention(bf16): pitches and batch strides must be multiples of 8 elements");
  TCD_REQUIRE(((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0, "tcd_attention(bf16): 16-byte pointer alignment");
  TCD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || rng_state), "tcd_attention(bf16): bad dropout arguments");
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap_3d_bf16(&tq, Q, (int64_t)heads * fa::HD, Lq, samples, ldq, qbs, fa::BQ))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, K, (int64_t)heads * fa::HD, Lk, samples, ldk, kbs, fa::BKV))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, V, (int64_t)heads * fa::HD, Lk, samples, ldv, vbs, fa::BKV))) return rc;
  if ((rc = make_tmap_3d_bf16(&to, O, (int64_t)heads * fa::HD, Lq, samples, ldo, obs, fa::BQ))) return rc;
  const int64_t items = (int64_t)ceil_div(Lq, fa::BQ) * heads * samples;
  const int resident = 2 * num_sms();                      // two CTAs per SM (smem / TMEM / registers)
  const int grid = (int)(items < resident ? items : resident);
  const float sl2 = scale * 1.4426950408889634f;
#if TCD_TUNE_ATTN_2Q
  if (Lq > fa::BQ) {                                       // at least two query tiles: the two-tile ping-pong kernel
    if (dropout_p > 0.f)
      return launch_attention_tc2q<true>(tq, tk, tv, to, Lq, Lk, heads, samples, sl2, lse, drop_threshold(dropout_p),
                                         1.0f / (1.0f - dropout_p), (const uint64_t*)rng_state, site, st);
    return launch_attention_tc2q<false>(tq, tk, tv, to, Lq, Lk, heads, samples, sl2, lse, 0u, 1.0f, nullptr, 0u, st);
  }
#endif
  if (dropout_p > 0.f)
    return launch_attention_tc<true>(tq, tk, tv, to, grid, Lq, Lk, heads, samples, sl2, lse, drop_threshold(dropout_p),
                                     1.0f / (1.0f - dropout_p), (const uint64_t*)rng_state, site, st);
  return launch_attention_tc<false>(tq, tk, tv, to, grid, Lq, Lk, heads, samples, sl2, lse, 0u, 1.0f, nullptr, 0u, st);
}

}  // namespace tcd
