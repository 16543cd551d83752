// bf16 flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 64, no mask:
//   O = softmax(scale * Q K^T) V   per (sample, head);  replaces SBI_MSA's core (model/model.py:97-102)
//   and the nn.MultiheadAttention core of the music encoder in bf16 mode.
//
// Persistent CTAs (2 per SM: ~97 KiB smem, 256 TMEM columns and <= 96 registers each) loop over work items = one
// 128-query tile of one (sample, head).  Keys are consumed 64 at a time:
//   warp 0      TMA producer: Q tile per item, then K(0) V(0) K(1) V(1) ... as 64-key x 64 bf16 boxes (3-D tensor
//               maps (col, row, sample): rows past Lq/Lk are zero-filled, never read from the next sample) into a
//               6-slot shared-memory ring that keeps prefetching across item boundaries
//   warp 1      MMA issuer (one lane).  S(t) = Q K(t)^T : tcgen05.mma M=128 N=kw K=16 x4 into TMEM buffer t%2 —
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
// before P V(t) 435 -> persistent CTAs (Lk=152: 211 -> 268) -> this version (S double-buffered, O in TMEM).
#include "tc_attn_common.cuh"
#include "dropout.cuh"

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
// exp2 of two arguments on the FMA pipe (packed f32x2 instructions), for the part of each tile that is taken off the
// MUFU pipe: x = n + f with n = round(x) (magic-number add), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial
// (relative error 7.5e-5, far below the bf16 rounding of P); 2^n by adding n to the exponent field.  Arguments are
// clamped at -126 (result ~1e-38, i.e. zero after the bf16 rounding); x <= 8 by the lazy reference max.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 fl = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(fl, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(make_float2(0.055171475f, 0.055171475f), f, make_float2(0.24261111f, 0.24261111f));
  p = __ffma2_rn(p, f, make_float2(0.69326103f, 0.69326103f));
  p = __ffma2_rn(p, f, make_float2(0.99992806f, 0.99992806f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23)));
}
// POLY = pairs of every 8 scores whose exp2 runs on the FMA pipe instead of the MUFU pipe (full tiles only).
template <bool MASKED, bool DROP, int POLY>
__device__ __forceinline__ float exp_store32(const uint32_t (&raw)[32], int valid, float scale_log2, float mt, uint32_t rowb,
                                             int chunk0, int r, uint32_t rowseed, uint32_t key0, uint32_t thr, float rk) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float p[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!MASKED && e < 2 * POLY) {
        if ((e & 1) == 0) {
          const float2 q = ex2_poly2(__ffma2_rn(make_float2(__uint_as_float(raw[8 * j + e]), __uint_as_float(raw[8 * j + e + 1])),
                                                make_float2(scale_log2, scale_log2), make_float2(-mt, -mt)));
          p[e] = q.x;
          p[e + 1] = q.y;
        }
        continue;
      }
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

// VAR (tuning variant, TCD_ATTN_VAR): bit 0 = the row-max / row-sum exchange synchronises only the two warps that
// share a row (named barriers 2..5, 64 threads) instead of all eight softmax warps; bit 1 = no wait on o_full before
// P(t) overwrites the buffer P V(t-2) read (s_full of S(t), already observed, was committed after P V(t-2) by the same
// thread, and tcgen05.commit covers every earlier MMA of that thread); bits 2-3 = POLY of exp_store32; bit 4 = role swap, bit 5 / bit 6 = converged MMA / producer issue loops, bit 7 = deferred epilogue (below).
template <bool DROP, int VAR>
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

  // `warp` is the ROLE index (0 producer, 1 MMA issuer, 2..9 softmax).  VAR bit 4 gives the two single-thread roles the
  // highest physical warp ids (8, 9): the warp scheduler favours high warp ids among eligible warps, and the
  // producer / MMA hand-offs are on the critical path of every tile.  Softmax warps then are physical warps 0..7
  // (their TMEM lane quarter always follows the physical warp id).
  constexpr bool ROLE_SWAP = (VAR & 16) != 0;
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = ROLE_SWAP ? (pwarp >= SM_WARPS ? pwarp - SM_WARPS : pwarp + 2) : pwarp;
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
    if constexpr ((VAR & 64) != 0) {                        // converged loop, one elected lane issues (see the MMA role)
      const uint32_t leader = elect_one();
      int slot = 0;
      uint32_t ph = 0, it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int q0 = (w % qtiles) * BQ, h = (w / qtiles) % heads, b = w / (qtiles * heads);
        mbar_wait(q_empty, (it & 1u) ^ 1u);
        mbar_expect_tx_p(leader, q_full, Q_BYTES);
        tma_load_3d_p(leader, sQ, &tm_q, q_full, h * HD, q0, b);
        for (int item = 0; item < 2 * nt; ++item) {
          mbar_wait(empty(slot), ph ^ 1u);
          mbar_expect_tx_p(leader, full(slot), KV_BYTES);
          tma_load_3d_p(leader, sRing + slot * KV_BYTES, (item & 1) ? &tm_v : &tm_k, full(slot), h * HD, (item >> 1) * BKV, b);
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
    } else if (lane == 0) {
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
    if constexpr ((VAR & 32) != 0) {
      // Converged issue loop (VAR bit 5).  With the loop under `if (lane == 0)` ptxas keeps every descriptor in vector
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
              tc_mma_p(leader, tmem + O_COL + ((VAR & 128) != 0 ? 64u * (it & 1u) : 0u), desc128(pbase + (uint32_t)(k * 32)),
                       desc128(vbase + (uint32_t)(k * 2048)), id_pv, (uint32_t)(t | k));
          tc_commit_p(leader, empty(vs));
          tc_commit_p(leader, o_full(b));
          vs += 2;
          if (vs > NSLOT) { vs = 1; vph ^= 1u; }
        }
      }
    } else
    if (lane == 0) {
      const uint64_t qdesc = desc128(sQ);
      auto kw_of = [&](int t) { int r = Lk - t * BKV; r = r < BKV ? r : BKV; return (r + 15) & ~15; };
      int g0 = 0;                                          // ring counter at the start of the current work item
      int tc0 = 0;                                         // KV-tile counter at the start of the current work item
      int it = 0;
      auto issue_s = [&](int t) {                          // S(t) -> TMEM buffer (tc0+t)&1 (free: its P V predecessor
        const int g = g0 + 2 * t, slot = g % NSLOT;        //  was issued only after the softmax released that buffer)
        mbar_wait(full(slot), (uint32_t)(g / NSLOT) & 1u);
        tc_fence_after();
        const uint64_t kdesc = desc128(sRing + slot * KV_BYTES);
        const uint32_t id = idesc(kw_of(t), 0);
        const uint32_t d = tmem + S_COL + 64u * (uint32_t)((tc0 + t) & 1);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma(d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), id, k != 0);
        tc_commit(empty(slot));
        if (t == nt - 1) tc_commit(q_empty);               // Q may be overwritten by the next work item
        tc_commit(s_full((tc0 + t) & 1));
      };
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it, g0 += 2 * nt, tc0 += nt) {
        mbar_wait(q_full, (uint32_t)it & 1u);
        tc_fence_after();
        issue_s(0);
        for (int t = 0; t < nt; ++t) {
          const int tc = tc0 + t, b = tc & 1;
          if (t + 1 < nt) issue_s(t + 1);                   // one tile ahead of the softmax
          mbar_wait(p_full(b), (uint32_t)(tc >> 1) & 1u);   // P(t) in smem, S(t) consumed, O rescaled if needed
          tc_fence_after();
          const int g = g0 + 2 * t + 1, slot = g % NSLOT;
          mbar_wait(full(slot), (uint32_t)(g / NSLOT) & 1u);
          tc_fence_after();
          const uint32_t vbase = sRing + slot * KV_BYTES, pbase = sP + (uint32_t)(b * P_BYTES);
          const uint32_t id = idesc(HD, 1);
          const int ksteps = kw_of(t) / 16;
          for (int k = 0; k < ksteps; ++k)                  // A = P (K-major, +32 B per 16 keys), B = V (MN-major, +2 KiB)
            tc_mma(tmem + O_COL, desc128(pbase + (uint32_t)(k * 32)), desc128(vbase + (uint32_t)(k * 2048)), id, (t | k) != 0);
          tc_commit(empty(slot));
          tc_commit(o_full(b));
        }
      }
    }
  } else {
    // ===================== softmax / output (8 warps, two threads per query row) =====================
    const int sw = warp - 2;
    const int quarter = pwarp & 3;                // TMEM lanes [32*quarter, +32) are visible to this (physical) warp
    const int hh = sw >> 2;                       // which 32-key half of the tile / 32-column half of the output
    const int r = quarter * 32 + lane;
    constexpr bool PAIR_BAR = (VAR & 1) != 0, SKIP_OWAIT = (VAR & 2) != 0;
    constexpr int POLY = (VAR >> 2) & 3;
    auto pair_sync = [&]() {
      if constexpr (PAIR_BAR) {
        if (quarter == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
        else if (quarter == 1) asm volatile("bar.sync 3, 64;" ::: "memory");
        else if (quarter == 2) asm volatile("bar.sync 4, 64;" ::: "memory");
        else asm volatile("bar.sync 5, 64;" ::: "memory");
      } else {
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    };
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem_gen + OFF_X);
    int tc = 0;                                            // KV-tile counter across work items
    uint32_t dseed = 0;
    if constexpr (DROP) dseed = drop_site_seed(rng_state, drop_site);
    // VAR bit 7 (needs bit 5): the output row accumulates in TMEM buffer O[item parity] and an item's epilogue is
    // DEFERRED until the first key tile of the CTA's next item has been processed, so the softmax warps never idle
    // waiting for the item's last P V (~1 200 cycles per item, profiles/r01_issue_loops.md).
    constexpr bool DEFER = (VAR & 128) != 0;
    // final O (TMEM) / l -> bf16 -> swizzled staging tile (a P buffer no P V is reading) -> TMA store.
    //   tcl = the item's last key tile (its P V must have retired); tnext = the next tile this CTA will process
    //   (its max-exchange buffer is idle; in DEFER mode its P buffer is the free one).
    auto finish = [&](int w_e, float m_e, float l_e, uint32_t ocol, int tcl, int tnext) {
      const int q0 = (w_e % qtiles) * BQ, h = (w_e / qtiles) % heads, b = w_e / (qtiles * heads);
      mbar_wait(o_full(tcl & 1), (uint32_t)(tcl >> 1) & 1u);
      tc_fence_after();
      uint32_t ov[32];
      tc_ld32(lane_addr + ocol + hh * 32, ov);
      tc_wait_ld();
      tc_fence_before();
      float* xs = xch + (tnext & 1) * 256;
      xs[r * 2 + hh] = l_e;
      pair_sync();
      const float lsum = xs[r * 2] + xs[r * 2 + 1];
      const float inv = 1.0f / lsum;
      // log2-domain log-sum-exp of the scaled scores (training: the backward pass recomputes P = exp2(s*c - lse))
      if (lse != nullptr && hh == 0 && q0 + r < Lq) lse[((int64_t)b * heads + h) * Lq + q0 + r] = m_e + log2f(lsum);
      const uint32_t stage = sP + (DEFER ? (uint32_t)((tnext & 1) * P_BYTES) : 0u);
      const uint32_t rowo = stage + (uint32_t)(r * 128);
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
        tma_store_3d(&tm_o, stage, h * HD, q0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tile is a P buffer of the next tiles
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    bool pending = false;                                  // DEFER: the previous item's epilogue is still to do
    int w_p = 0;
    float m_p = 0.f, l_p = 0.f;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int q0 = (w % qtiles) * BQ, h = (w / qtiles) % heads, b = w / (qtiles * heads);
      const uint32_t rowseed = dseed ^ (uint32_t)((b * heads + h) * Lq + q0 + r) * kDropC1;
      const uint32_t ocol = O_COL + (DEFER ? 64u * (it & 1u) : 0u);
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
          tc_ld32(lane_addr + ocol + hh * 32, ov);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * corr);
          tc_st32(lane_addr + ocol + hh * 32, ov);
          tc_wait_st();
        }
        l *= corr;
        m = mt;
        if (!SKIP_OWAIT && tc >= 2) mbar_wait(o_full(sb), (uint32_t)((tc - 2) >> 1) & 1u);   // P V(tc-2) has finished reading P buffer sb
        // p = exp2(s*scale - m), partial row sum, bf16 P(t) into swizzled smem (row r, 16-byte chunk j at j ^ (r & 7))
        if (valid > 0) {
          const uint32_t rowb = sP + (uint32_t)(sb * P_BYTES + r * 128);
          const uint32_t key0 = (uint32_t)(t * BKV + hh * 32);
          l += valid >= 32 ? exp_store32<false, DROP, POLY>(raw, 32, scale_log2, mt, rowb, hh * 4, r, rowseed, key0, drop_thr, drop_rk)
                           : exp_store32<true, DROP, 0>(raw, valid, scale_log2, mt, rowb, hh * 4, r, rowseed, key0, drop_thr, drop_rk);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(sb));
        if (DEFER && t == 0 && pending)                        // previous item: its last tile was tc - 1, next tile is tc + 1
          finish(w_p, m_p, l_p, O_COL + 64u * ((it - 1u) & 1u), tc - 1, tc + 1);
      }
      if constexpr (DEFER) {
        pending = true; w_p = w; m_p = m; l_p = l;
      } else {
        finish(w, m, l, O_COL, tc - 1, tc);
      }
    }  // work items
    if (DEFER && pending) finish(w_p, m_p, l_p, O_COL + 64u * ((it - 1u) & 1u), tc - 1, tc);
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

template <bool DROP, int VAR>
static int launch_attention_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int grid,
                               int Lq, int Lk, int heads, int samples, float scale_log2, float* lse, uint32_t thr, float rk,
                               const uint64_t* rng_state, uint32_t site, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fa::attention_tc_kernel<DROP, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa::SMEM);
    if (e != cudaSuccess) { set_error("attention_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  fa::attention_tc_kernel<DROP, VAR><<<grid, fa::THREADS, fa::SMEM, st>>>(tq, tk, tv, to, Lq, Lk, heads, samples, scale_log2, lse, thr,
                                                                         rk, rng_state, site);
  return check_launch("attention_tc");
}

// Tuning variant of the kernel (see the VAR comment above).  TCD_ATTN_VAR overrides the default for A/B measurements
// (tools/kernel_bench.py attn); every variant computes the same function and passes the same parity tests.
constexpr int kAttnDefaultVar = 35;   // r01 A/B (profiles/r01_issue_loops.md): 0.3175 -> 0.2949 ms self, 0.1065 -> 0.1004 ms cross.
// The exp2-polynomial variants (bits 2-3) are NOT adopted: +0.7 % at best, and VAR 39 fails the training gradient test
// with dropout (tests/test_gpu_train.py::test_training_gradients_with_dropout_vs_oracle: whole-gradient cosine 0.935).
static int attention_variant() {
  static int var = -1;
  if (var < 0) {
    const char* e = getenv("TCD_ATTN_VAR");
    int v = e ? atoi(e) : kAttnDefaultVar;
    if (v != 0 && v != 3 && v != 7 && v != 11 && v != 19 && v != 23 && v != 35 && v != 39 && v != 43 && v != 163 && v != 99 && v != 103) v = kAttnDefaultVar;
    var = v;
  }
  return var;
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
#define TCD_ATTN_LAUNCH(VARV)                                                                                              \
  (dropout_p > 0.f ? launch_attention_tc<true, VARV>(tq, tk, tv, to, grid, Lq, Lk, heads, samples, sl2, lse,                \
                                                     drop_threshold(dropout_p), 1.0f / (1.0f - dropout_p),                 \
                                                     (const uint64_t*)rng_state, site, st)                                 \
                   : launch_attention_tc<false, VARV>(tq, tk, tv, to, grid, Lq, Lk, heads, samples, sl2, lse, 0u, 1.0f,     \
                                                      nullptr, 0u, st))
  switch (attention_variant()) {
    case 3: return TCD_ATTN_LAUNCH(3);
    case 7: return TCD_ATTN_LAUNCH(7);
    case 11: return TCD_ATTN_LAUNCH(11);
    case 19: return TCD_ATTN_LAUNCH(19);
    case 23: return TCD_ATTN_LAUNCH(23);
    case 39: return TCD_ATTN_LAUNCH(39);
    case 43: return TCD_ATTN_LAUNCH(43);
    case 163: return TCD_ATTN_LAUNCH(163);
    case 99: return TCD_ATTN_LAUNCH(99);
    case 103: return TCD_ATTN_LAUNCH(103);
    case 0: return TCD_ATTN_LAUNCH(0);
    default: return TCD_ATTN_LAUNCH(35);
  }
#undef TCD_ATTN_LAUNCH
}

}  // namespace tcd
