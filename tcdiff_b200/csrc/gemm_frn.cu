// Fused block tail: the `fc` / `linear2` GEMM with the FiLM + residual + LayerNorm tail in its epilogue.
//
//   y      = A W^T (+ bias)                         A (M,K) bf16, W (512,K) bf16 (nn.Linear weight as stored)
//   z      = LN_in(y)            (optional; SBI_MSA.layer_norm, eps 1e-6, model/model.py:68,106)
//   x_out  = x_in + (1 + scale) * z + shift         (featurewise_affine + residual, model/model.py:171-173,327,334,339)
//   plain  = LN_next(x_out) -> bf16,  rot = rotary(plain) -> bf16   (operands of the next block)
//
// i.e. tcd_gemm followed by tcd_film_residual_norm without the bf16 round trip of y through HBM (2 + 2 of the
// 16 B/element the pair moves) and without the second kernel.  EXPERIMENTAL in round 1: the engine keeps the unfused
// pair unless TCD_FUSE_TAILS selects this kernel.
//
// One CTA per SM, persistent over 128-row tiles; the accumulator is the FULL 512-column row block, i.e. all 512
// TMEM columns (no accumulator double buffering: the epilogue of a tile and the MMAs of the next one do not overlap
// inside an SM, other SMs fill the memory pipes meanwhile).
//   warp 0      TMA producer: per 64-wide k block A 128x64 and W 512x64 (two 256-row boxes) bf16, 128B swizzle,
//               2-stage ring of 80 KB
//   warp 1      MMA issuer (converged loop, elected lane): two tcgen05.mma M=128 N=256 K=16 per k step (column
//               halves of the row block), tcgen05.commit per stage / per tile
//   warps 2..9  epilogue, two threads per row (256 columns each; row statistics are exchanged through shared
//               memory between the two warps that share a TMEM lane quarter):
//               pass 1  LN_in statistics of y straight from TMEM (two-pass mean / variance)
//               pass 2  per 32-column chunk: x_in arrives by TMA in a per-warp 2-slot ring (32 rows x 128 B,
//                       swizzled: conflict-free row reads; a slot is refilled as soon as it has been read),
//                       v = x + (1+scale) z + shift replaces y in TMEM (tcgen05.st)
//               pass 2b variance of v from TMEM; the same sweep stages v in the idle x slots (alternating) and
//                       TMA-stores it to x_out
//               pass 3  per 64 columns: normalise, (rotate with a cos / sin tile staged through a slot,) pack to
//                       bf16, stage, TMA-store to plain / rot
// Row tails (M % 128) are zero-filled on load and clipped on store by the tensor maps.
#include <stdlib.h>

#include "tc_gemm_common.cuh"

namespace tcd {

int num_sms();
int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32);

namespace gf {

constexpr int FN = 512;                                  // the full row block
constexpr int FSTAGES = 2;
constexpr int FA_BYTES = BM * BK * 2;                    // 16 KB
constexpr int FWH_BYTES = 256 * BK * 2;                  // 32 KB: one 256-row box of W
constexpr int FSTAGE = FA_BYTES + 2 * FWH_BYTES;         // 80 KB
constexpr int SLOT = 32 * 128;                           // 32 rows x 128 bytes
constexpr int OFF_EPI = FSTAGES * FSTAGE;                // 160 KB
constexpr int OFF_XCH = OFF_EPI + EPI_WARPS * 2 * SLOT;  // + 64 KB
constexpr int OFF_BAR = OFF_XCH + 2 * 256 * 4;           // two exchange buffers of [128 rows][2 halves] floats
constexpr size_t SMEM = OFF_BAR + 256;                   // 231 680 B <= 232 448 (no alignment slack: checked below)

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
  const float* film;        // (samples, film_ld): [scale(512) | shift(512)] at film_off
  int64_t film_ld, film_off;
  const float* gnext;       // LN_next gamma / beta
  const float* bnext;
  float eps_next;
  const float* rot_cos;     // (tokens_per_sample, 256) or NULL
  const float* rot_sin;
  int has_xout, has_plain, has_rot;
  int M, K, tps;
  unsigned long long* dbg;  // optional 8 cycle counters (tcd_gemm_frn_set_debug): epilogue warp 2 lane 0 / MMA warp of every CTA
};

template <bool DBG>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_frn_kernel(
    const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
    const __grid_constant__ CUtensorMap tm_xin, const __grid_constant__ CUtensorMap tm_xout,
    const __grid_constant__ CUtensorMap tm_plain, const __grid_constant__ CUtensorMap tm_rot, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();                    // SWIZZLE_128B atoms need 1024-byte alignment
  auto full_bar = [&](int s) { return base + OFF_BAR + 8u * s; };
  auto empty_bar = [&](int s) { return base + OFF_BAR + 8u * (FSTAGES + s); };
  const uint32_t tfull = base + OFF_BAR + 8u * (2 * FSTAGES), tempty = tfull + 8u;
  auto x_bar = [&](int ew, int s) { return base + OFF_BAR + 8u * (2 * FSTAGES + 2 + ew * 2 + s); };
  const uint32_t tmem_slot = base + OFF_BAR + 8u * (2 * FSTAGES + 2 + 2 * EPI_WARPS);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + 8 * (2 * FSTAGES + 2 + 2 * EPI_WARPS));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.M;
  const int num_tiles = (M + BM - 1) / BM;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_xin) : "memory");
    for (int s = 0; s < FSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, EPI_WARPS);
    for (int e = 0; e < EPI_WARPS; ++e) { mbar_init(x_bar(e, 0), 1); mbar_init(x_bar(e, 1), 1); }
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
    // ===================== TMA producer (A / W ring) =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = tile * BM;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx_p(leader, full_bar(stage), FSTAGE);
        const uint32_t sa = base + stage * FSTAGE;
        tma_load_2d_p(leader, sa, &tm_a, full_bar(stage), kb * BK, m0);
        tma_load_2d_p(leader, sa + FA_BYTES, &tm_w, full_bar(stage), kb * BK, 0);
        tma_load_2d_p(leader, sa + FA_BYTES + FWH_BYTES, &tm_w, full_bar(stage), kb * BK, 256);
        if (++stage == FSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    long long d_tempty = 0, d_loop = 0, d_full = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const long long c0 = tick<DBG>();
      mbar_wait(tempty, ((uint32_t)it & 1u) ^ 1u);          // the epilogue has drained the row block
      tc_fence_after();
      const long long c1 = tick<DBG>();
      d_tempty += c1 - c0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const long long f0 = tick<DBG>();
        mbar_wait(full_bar(stage), phase);
        d_full += tick<DBG>() - f0;
        tc_fence_after();
        const uint32_t sa = base + stage * FSTAGE;
        const uint64_t adesc = umma_desc_k128(sa);
        const uint64_t b0 = umma_desc_k128(sa + FA_BYTES), b1 = umma_desc_k128(sa + FA_BYTES + FWH_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          tc_mma_f16_p(leader, tmem_base, adesc + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), kIdesc, (uint32_t)(kb | k));
          tc_mma_f16_p(leader, tmem_base + 256u, adesc + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), kIdesc, (uint32_t)(kb | k));
        }
        tc_commit_p(leader, empty_bar(stage));
        if (++stage == FSTAGES) { stage = 0; phase ^= 1u; }
      }
      tc_commit_p(leader, tfull);
      d_loop += tick<DBG>() - c1;
    }
    if (DBG && p.dbg != nullptr && lane == 0) {
      atomicAdd(p.dbg + 5, (unsigned long long)d_tempty);
      atomicAdd(p.dbg + 6, (unsigned long long)d_loop);
      atomicAdd(p.dbg + 7, (unsigned long long)d_full);
    }
  } else {
    // ===================== epilogue (8 warps, two threads per row) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                          // TMEM lanes [32*quarter, +32) are visible to this warp
    const int half = ew >> 2;                              // columns [256*half, +256)
    const int r = quarter * 32 + lane;                     // row inside the tile
    const int cb = half * 256;
    const uint32_t slot0 = base + OFF_EPI + (uint32_t)(ew * 2 * SLOT), slot1 = slot0 + SLOT;
    const uint32_t xb0 = x_bar(ew, 0), xb1 = x_bar(ew, 1);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cb;
    float* xch = reinterpret_cast<float*>(smem_raw + OFF_XCH);
    int xk = 0;
    auto pair_sum = [&](float v) -> float {                // sum over the two threads that share row r
      float* xs = xch + (xk & 1) * 256;
      ++xk;
      xs[r * 2 + half] = v;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
      return xs[r * 2] + xs[r * 2 + 1];
    };
    auto issue_x = [&](int m0, int c) {                    // lane 0: x_in rows [m0 + 32*quarter, +32), columns [cb + 32c, +32)
      const uint32_t bar = (c & 1) ? xb1 : xb0;
      mbar_expect_tx(bar, SLOT);
      tma_load_2d((c & 1) ? slot1 : slot0, &tm_xin, bar, cb + 32 * c, m0 + quarter * 32);
    };
    auto load_y = [&](int c, float (&yv)[32]) {            // y chunk c of this thread's row (+ bias)
      uint32_t raw[32];
      tc_ld32(lane_taddr + (uint32_t)(32 * c), raw);
      tc_wait_ld();
      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + cb + 32 * c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j);
          yv[4 * j] = __uint_as_float(raw[4 * j]) + b.x;
          yv[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b.y;
          yv[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b.z;
          yv[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) yv[j] = __uint_as_float(raw[j]);
      }
    };
    constexpr float invD = 1.0f / (float)FN;

    if (blockIdx.x < num_tiles && lane == 0) {
      issue_x(blockIdx.x * BM, 0);
      issue_x(blockIdx.x * BM, 1);
    }
    int it = 0;
    long long d_e[5] = {0, 0, 0, 0, 0};                    // cycles: wait for the MMAs, pass 1, 2, 2b, 3
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = tile * BM;
      const int row0 = m0 + quarter * 32;
      const bool live = row0 < M;                          // warp-uniform: some row of this warp exists
      const int grow = min(m0 + r, M - 1);                 // clamped: rows past M compute garbage that is clipped on store
      const float* fp = p.film + (int64_t)(grow / p.tps) * p.film_ld + p.film_off;
      const long long e0 = tick<DBG>();
      mbar_wait(tfull, (uint32_t)it & 1u);
      tc_fence_after();
      const long long e1 = tick<DBG>();

      // ---- pass 1: statistics of y for the inner LayerNorm
      float mean1 = 0.f, rstd1 = 1.f;
      if (p.gin != nullptr) {
        float s = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float yv[32];
          load_y(c, yv);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) { s0 += yv[j]; s1 += yv[j + 1]; s2 += yv[j + 2]; s3 += yv[j + 3]; }
          s += (s0 + s1) + (s2 + s3);
        }
        mean1 = pair_sum(s) * invD;
        float q = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float yv[32];
          load_y(c, yv);
          float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float a = yv[j] - mean1, b = yv[j + 1] - mean1, cc = yv[j + 2] - mean1, d = yv[j + 3] - mean1;
            q0 += a * a; q1 += b * b; q2 += cc * cc; q3 += d * d;
          }
          q += (q0 + q1) + (q2 + q3);
        }
        rstd1 = 1.0f / sqrtf(pair_sum(q) * invD + p.eps_in);
      }

      const long long e2 = tick<DBG>();
      // ---- pass 2: v = x + (1 + scale) * z + shift, chunk by chunk, into TMEM (replacing y).  The x slots are only
      //      read here, so a slot is refilled (chunk c + 2) as soon as every lane has read chunk c: true double buffering.
      float sv = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const uint32_t slot = (c & 1) ? slot1 : slot0;
        mbar_wait((c & 1) ? xb1 : xb0, (uint32_t)(c >> 1) & 1u);
        float yv[32];
        load_y(c, yv);
        const int col = cb + 32 * c;
        if (p.gin != nullptr) {
          const float4* gp = reinterpret_cast<const float4*>(p.gin + col);
          const float4* bp = reinterpret_cast<const float4*>(p.bin + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g = __ldg(gp + j), b = __ldg(bp + j);
            yv[4 * j] = (yv[4 * j] - mean1) * rstd1 * g.x + b.x;
            yv[4 * j + 1] = (yv[4 * j + 1] - mean1) * rstd1 * g.y + b.y;
            yv[4 * j + 2] = (yv[4 * j + 2] - mean1) * rstd1 * g.z + b.z;
            yv[4 * j + 3] = (yv[4 * j + 3] - mean1) * rstd1 * g.w + b.w;
          }
        }
        const uint32_t rb = slot + (uint32_t)(lane * 128);
        const float4* scp = reinterpret_cast<const float4*>(fp + col);
        const float4* shp = reinterpret_cast<const float4*>(fp + FN + col);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        uint32_t vr[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 xv = lds128(rb + (uint32_t)(((j ^ lane) & 7) << 4));   // SWIZZLE_128B: 16-byte chunk j of row `lane`
          const float4 sc = __ldg(scp + j), sh = __ldg(shp + j);
          const float v0 = xv.x + ((sc.x + 1.0f) * yv[4 * j] + sh.x);
          const float v1 = xv.y + ((sc.y + 1.0f) * yv[4 * j + 1] + sh.y);
          const float v2 = xv.z + ((sc.z + 1.0f) * yv[4 * j + 2] + sh.z);
          const float v3 = xv.w + ((sc.w + 1.0f) * yv[4 * j + 3] + sh.w);
          a0 += v0; a1 += v1; a2 += v2; a3 += v3;
          vr[4 * j] = __float_as_uint(v0); vr[4 * j + 1] = __float_as_uint(v1);
          vr[4 * j + 2] = __float_as_uint(v2); vr[4 * j + 3] = __float_as_uint(v3);
        }
        sv += (a0 + a1) + (a2 + a3);
        __syncwarp();                                      // every lane has read the slot
        if (c + 2 < 8 && lane == 0) issue_x(m0, c + 2);
        tc_st32(lane_taddr + (uint32_t)(32 * c), vr);
      }
      tc_wait_st();
      const long long e3 = tick<DBG>();

      // ---- pass 2b: variance of v from TMEM; the same sweep stages v in the (now idle) x slots, alternating, and
      //      sends it to x_out with one TMA store per 32-column chunk
      const float mean2 = pair_sum(sv) * invD;
      float q = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t raw[32];
        tc_ld32(lane_taddr + (uint32_t)(32 * c), raw);
        tc_wait_ld();
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float a = __uint_as_float(raw[j]) - mean2, b = __uint_as_float(raw[j + 1]) - mean2;
          const float cc = __uint_as_float(raw[j + 2]) - mean2, d = __uint_as_float(raw[j + 3]) - mean2;
          q0 += a * a; q1 += b * b; q2 += cc * cc; q3 += d * d;
        }
        q += (q0 + q1) + (q2 + q3);
        if (p.has_xout) {
          const uint32_t slot = (c & 1) ? slot1 : slot0;
          if (c >= 2) {                                    // the store of chunk c-2 has read this slot (c-1 may be pending)
            if (lane == 0) tma_store_wait_read1();
            __syncwarp();
          }
          const uint32_t rb = slot + (uint32_t)(lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(rb + (uint32_t)(((j ^ lane) & 7) << 4), raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && live) {
            tma_store_2d(&tm_xout, slot, cb + 32 * c, row0);
            tma_store_commit();
          }
        }
      }
      const float rstd2 = 1.0f / sqrtf(pair_sum(q) * invD + p.eps_next);
      const long long e4 = tick<DBG>();

      // ---- pass 3: LayerNorm of v -> bf16 operands of the next block, 64 columns (one 128-byte bf16 row) per iteration.
      //      plain only: output staging alternates between the two slots (no wait on the store just issued);
      //      rot: slot0 stages this warp's 32 x 32 cos / sin tile (coalesced loads, each thread then reads its own
      //      row), slot1 stages the output.
      const bool single_plain = p.has_plain && !p.has_rot;
      if (!single_plain) {                                 // slot0 / slot1 get fixed roles: every x_out store must be done
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll 1
      for (int c2 = 0; c2 < 4; ++c2) {
        const int col = cb + 64 * c2;
        float nv[64];
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          uint32_t raw[32];
          tc_ld32(lane_taddr + (uint32_t)(64 * c2 + 32 * hq), raw);
          tc_wait_ld();
          const float4* gp = reinterpret_cast<const float4*>(p.gnext + col + 32 * hq);
          const float4* bp = reinterpret_cast<const float4*>(p.bnext + col + 32 * hq);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g = __ldg(gp + j), b = __ldg(bp + j);
            nv[32 * hq + 4 * j] = (__uint_as_float(raw[4 * j]) - mean2) * rstd2 * g.x + b.x;
            nv[32 * hq + 4 * j + 1] = (__uint_as_float(raw[4 * j + 1]) - mean2) * rstd2 * g.y + b.y;
            nv[32 * hq + 4 * j + 2] = (__uint_as_float(raw[4 * j + 2]) - mean2) * rstd2 * g.z + b.z;
            nv[32 * hq + 4 * j + 3] = (__uint_as_float(raw[4 * j + 3]) - mean2) * rstd2 * g.w + b.w;
          }
        }
        if (single_plain) {
          const uint32_t slot = (c2 & 1) ? slot1 : slot0;
          if (lane == 0) tma_store_wait_read1();           // the store that last read this slot is done (the latest may be pending)
          __syncwarp();
          const uint32_t rb = slot + (uint32_t)(lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(rb + (uint32_t)(((j ^ lane) & 7) << 4), bf2(nv[8 * j], nv[8 * j + 1]), bf2(nv[8 * j + 2], nv[8 * j + 3]),
                   bf2(nv[8 * j + 4], nv[8 * j + 5]), bf2(nv[8 * j + 6], nv[8 * j + 7]));
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && live) {
            tma_store_2d(&tm_plain, slot, col, row0);
            tma_store_commit();
          }
        } else {
          const uint32_t rb1 = slot1 + (uint32_t)(lane * 128);
          if (p.has_plain) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(rb1 + (uint32_t)(((j ^ lane) & 7) << 4), bf2(nv[8 * j], nv[8 * j + 1]), bf2(nv[8 * j + 2], nv[8 * j + 3]),
                     bf2(nv[8 * j + 4], nv[8 * j + 5]), bf2(nv[8 * j + 6], nv[8 * j + 7]));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && live) {
              tma_store_2d(&tm_plain, slot1, col, row0);
              tma_store_commit();
            }
          }
          if (p.has_rot) {
            // interleaved pairs (2i, 2i+1) rotate by angle i of the token's table row (rotary_embedding_torch.py:107-113).
            // The warp's 32 rows x 32 angles of cos, then sin, pass through slot0: lane l loads the 16-byte chunk
            // (l & 7) of rows (l >> 3) + 4k (coalesced 128-byte rows), each thread reads back its own row.
            float cs[32], sn[32];
            const uint32_t rb0 = slot0 + (uint32_t)(lane * 128);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const float* tab = t == 0 ? p.rot_cos : p.rot_sin;
              float4 g4[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int ri = (lane >> 3) + 4 * k;
                const int pr = min(row0 + ri, M - 1) % p.tps;
                g4[k] = __ldg(reinterpret_cast<const float4*>(tab + (int64_t)pr * (FN / 2) + col / 2) + (lane & 7));
              }
              __syncwarp();                                // the previous read-back of slot0 is complete
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int ri = (lane >> 3) + 4 * k;
                sts128(slot0 + (uint32_t)(ri * 128) + (uint32_t)((((lane & 7) ^ ri) & 7) << 4), __float_as_uint(g4[k].x),
                       __float_as_uint(g4[k].y), __float_as_uint(g4[k].z), __float_as_uint(g4[k].w));
              }
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v4 = lds128(rb0 + (uint32_t)(((j ^ lane) & 7) << 4));
                if (t == 0) { cs[4 * j] = v4.x; cs[4 * j + 1] = v4.y; cs[4 * j + 2] = v4.z; cs[4 * j + 3] = v4.w; }
                else { sn[4 * j] = v4.x; sn[4 * j + 1] = v4.y; sn[4 * j + 2] = v4.z; sn[4 * j + 3] = v4.w; }
              }
            }
            if (lane == 0) tma_store_wait_read();          // the previous store from slot1 (issued an iteration ago) is done
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float* n8 = nv + 8 * j;
              const float* c4 = cs + 4 * j;
              const float* s4 = sn + 4 * j;
              sts128(rb1 + (uint32_t)(((j ^ lane) & 7) << 4),
                     bf2(n8[0] * c4[0] - n8[1] * s4[0], n8[1] * c4[0] + n8[0] * s4[0]),
                     bf2(n8[2] * c4[1] - n8[3] * s4[1], n8[3] * c4[1] + n8[2] * s4[1]),
                     bf2(n8[4] * c4[2] - n8[5] * s4[2], n8[5] * c4[2] + n8[4] * s4[2]),
                     bf2(n8[6] * c4[3] - n8[7] * s4[3], n8[7] * c4[3] + n8[6] * s4[3]));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && live) {
              tma_store_2d(&tm_rot, slot1, col, row0);
              tma_store_commit();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      d_e[0] += e1 - e0; d_e[1] += e2 - e1; d_e[2] += e3 - e2; d_e[3] += e4 - e3; d_e[4] += tick<DBG>() - e4;
      if (lane == 0) {
        mbar_arrive(tempty);                               // the MMAs of the next tile may overwrite the row block
        const int next = tile + gridDim.x;
        if (next < num_tiles) {                            // x chunks 0 / 1 of the next tile fly during its MMAs
          tma_store_wait_read();
          issue_x(next * BM, 0);
          issue_x(next * BM, 1);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
    if (DBG && p.dbg != nullptr && ew == 0 && lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) atomicAdd(p.dbg + k, (unsigned long long)d_e[k]);
    }
  }
  tc_fence_before();
  __syncthreads();
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
                  void* out_rot, const float* rot_cos, const float* rot_sin, int tps, cudaStream_t st) {
  CUtensorMap ta, tw, txi, txo, tp, tr;
  int rc = make_tmap_2d(&ta, A, M, K, lda, BM, false);
  if (rc) return rc;
  rc = make_tmap_2d(&tw, W, gf::FN, K, ldw, 256, false);
  if (rc) return rc;
  rc = make_tmap_2d(&txi, x_in, M, gf::FN, gf::FN, 32, true);
  if (rc) return rc;
  txo = txi;
  if (x_out) { rc = make_tmap_2d(&txo, x_out, M, gf::FN, gf::FN, 32, true); if (rc) return rc; }
  tp = ta;
  if (out_plain) { rc = make_tmap_2d(&tp, out_plain, M, gf::FN, gf::FN, 32, false); if (rc) return rc; }
  tr = ta;
  if (out_rot) { rc = make_tmap_2d(&tr, out_rot, M, gf::FN, gf::FN, 32, false); if (rc) return rc; }
  gf::Params p;
  p.bias = bias; p.gin = gin; p.bin = bin; p.eps_in = eps_in;
  p.film = film; p.film_ld = film_ld; p.film_off = film_off;
  p.gnext = gnext; p.bnext = bnext; p.eps_next = eps_next;
  p.rot_cos = rot_cos; p.rot_sin = rot_sin;
  p.has_xout = x_out != nullptr; p.has_plain = out_plain != nullptr; p.has_rot = out_rot != nullptr;
  p.M = (int)M; p.K = (int)K; p.tps = tps;
  p.dbg = g_frn_dbg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gf::gemm_frn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gf::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gf::gemm_frn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gf::SMEM);
    if (e != cudaSuccess) { set_error("gemm_frn: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int tiles = (int)((M + BM - 1) / BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (p.dbg != nullptr)
    gf::gemm_frn_kernel<true><<<grid, GEMM_THREADS, gf::SMEM, st>>>(ta, tw, txi, txo, tp, tr, p);
  else
    gf::gemm_frn_kernel<false><<<grid, GEMM_THREADS, gf::SMEM, st>>>(ta, tw, txi, txo, tp, tr, p);
  return check_launch("gemm_frn");
}

}  // namespace tcd

// Profiling aid: buf = 8 device uint64 counters that every later launch of the fused kernel adds its phase cycle
// counts to (epilogue warp 2: wait for MMAs, pass 1, 2, 2b, 3; MMA warp: wait for the epilogue, main loop, of which
// waiting for TMA), or NULL to switch it off.
extern "C" int tcd_gemm_frn_set_debug(void* buf) {
  tcd::g_frn_dbg = reinterpret_cast<unsigned long long*>(buf);
  return TCD_OK;
}

extern "C" int tcd_gemm_film_residual_norm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                           int64_t M, int64_t K, const float* x_in, float* x_out,
                                           const float* ln_in_gamma, const float* ln_in_beta, float ln_in_eps,
                                           const float* film, int64_t film_ld, int64_t film_off,
                                           const float* next_gamma, const float* next_beta, float next_eps,
                                           void* out_plain, void* out_rot, const float* rot_cos, const float* rot_sin,
                                           int tokens_per_sample, void* stream) {
  using namespace tcd;
  TCD_REQUIRE(A && W && x_in && film && next_gamma && next_beta, "tcd_gemm_film_residual_norm: null pointer");
  TCD_REQUIRE((ln_in_gamma == nullptr) == (ln_in_beta == nullptr), "tcd_gemm_film_residual_norm: inner LN params");
  TCD_REQUIRE(out_plain || out_rot, "tcd_gemm_film_residual_norm: no output operand requested");
  TCD_REQUIRE(!out_rot || (rot_cos && rot_sin), "tcd_gemm_film_residual_norm: rotary table missing");
  TCD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && lda % 8 == 0 && ldw % 8 == 0,
              "tcd_gemm_film_residual_norm: A/W base and pitch must be 16-byte aligned");
  TCD_REQUIRE(((uintptr_t)x_in % 16 == 0) && ((uintptr_t)x_out % 16 == 0) && ((uintptr_t)out_plain % 16 == 0) &&
                  ((uintptr_t)out_rot % 16 == 0) && ((uintptr_t)film % 16 == 0) && ((uintptr_t)bias % 16 == 0),
              "tcd_gemm_film_residual_norm: 16-byte alignment");
  TCD_REQUIRE(film_ld % 4 == 0 && film_off % 4 == 0, "tcd_gemm_film_residual_norm: film alignment");
  TCD_REQUIRE(tokens_per_sample > 0 && M < (1LL << 31) && K > 0 && K < (1LL << 31), "tcd_gemm_film_residual_norm: bad shape");
  if (M == 0) return TCD_OK;
  return gemm_frn_bf16(A, lda, W, ldw, bias, M, K, x_in, x_out, ln_in_gamma, ln_in_beta, ln_in_eps, film, film_ld, film_off,
                       next_gamma, next_beta, next_eps, out_plain, out_rot, rot_cos, rot_sin, tokens_per_sample,
                       as_stream(stream));
}
