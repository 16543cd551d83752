// bf16 flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 64, no mask:
//   O = softmax(scale * Q K^T) V   per (sample, head);  replaces SBI_MSA's core (model/model.py:97-102)
//   and the nn.MultiheadAttention core of the music encoder in bf16 mode.
//
// CTA = one 128-query tile of one (sample, head); 6 warps; two CTAs co-reside per SM (96 KiB smem and 256
// TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//   warp 0      TMA producer: Q tile once, then K(0) V(0) K(1) V(1) ... 128-key x 64 bf16 boxes (3-D tensor
//               maps (col, row, sample): rows past Lq/Lk are zero-filled, never read from the next sample)
//               into a 3-slot shared-memory ring (full/empty mbarriers)
//   warp 1      MMA issuer (one lane):  S = Q K^T     tcgen05.mma M=128 N=kw K=16 x4, both operands K-major
//                                       O_t = P V     tcgen05.mma M=128 N=64 K=16 x kw/16, V is MN-major
//               accumulators in TMEM: S in columns [0,128), O_t in [128,192)
//   warps 2..9  softmax: two threads per query row (TMEM lane), one per 64-key half of the tile.  Two
//               passes over the thread's 64 S columns with tcgen05.ld (row max -> exchanged with the
//               partner through smem; then p = exp2(s*scale*log2e - m)), P written as bf16 into
//               128B-swizzled smem (the A operand of the P V product), running (m, l_half) and the
//               thread's 32 output columns kept in fp32 registers and rescaled FA2-style:
//               o = o*corr + O_t.  Final o/l goes through smem and one TMA bulk store (rows past Lq are
//               clipped by the tensor map).  (r01: one thread per row issued ~1400 instructions per tile
//               with 2 warps/scheduler and ran latency-bound at 306 TFLOP/s.)
// Keys past Lk in the last tile are masked to -inf; the last tile's MMA width kw is rounded up to 32 keys
// only (cross-attention: Lk = 152 = 128 + 24 -> second tile costs N=32, not N=128).
#include <cuda.h>

#include "common.cuh"

namespace tcd {

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld,
                      int64_t batch_stride, int box_rows);

namespace fa {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * 128;           // 128 rows x 64 bf16
constexpr int NSLOT = 3;
constexpr int SM_WARPS = 8;
constexpr int THREADS = (2 + SM_WARPS) * 32;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128;
// smem: Q | ring[3] | P (2 x 64-key blocks; block 0 doubles as the output staging tile) | barriers
constexpr int OFF_Q = 0, OFF_RING = TILE_BYTES, OFF_P = OFF_RING + NSLOT * TILE_BYTES, OFF_X = OFF_P + 2 * TILE_BYTES;
constexpr int OFF_BAR = OFF_X + 128 * 2 * 4;   // OFF_X: per-row exchange of the two half-row maxima / sums
constexpr size_t SMEM = 1024 + OFF_BAR + 128;

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
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// SWIZZLE_128B descriptor over [rows][128 B] tiles (8-row atoms of 1024 B).  The same field values describe the
// K-major operands (Q, K, P: 64 contiguous K elements per row) and the MN-major operand V (64 contiguous N
// elements per row, K = key index advancing by rows); the major-ness lives in the instruction descriptor.
__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16: D f32, A/B bf16, M=128; N and B-major-ness supplied
__device__ __forceinline__ uint32_t idesc(int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS, 2) attention_tc_kernel(
    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, int Lq, int Lk, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base + OFF_Q, sRing = base + OFF_RING, sP = base + OFF_P, bar = base + OFF_BAR;
  const uint32_t q_full = bar, s_full = bar + 8, p_full = bar + 16, o_full = bar + 24;
  auto full = [&](int i) { return bar + 32u + 8u * i; };
  auto empty = [&](int i) { return bar + 32u + 8u * (NSLOT + i); };
  const uint32_t tmem_slot = bar + 32u + 8u * 2 * NSLOT;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (base - smem_u32(smem_raw)) + OFF_BAR + 32 + 8 * 2 * NSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (Lk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o) : "memory");
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, SM_WARPS);
    mbar_init(o_full, 1);
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
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tm_q, q_full, h * HD, q0, b);
      for (int item = 0; item < 2 * nt; ++item) {          // K(0) V(0) K(1) V(1) ...
        const int slot = item % NSLOT;
        const uint32_t ph = (uint32_t)(item / NSLOT) & 1u;
        mbar_wait(empty(slot), ph ^ 1u);
        mbar_expect_tx(full(slot), TILE_BYTES);
        tma_load_3d(sRing + slot * TILE_BYTES, (item & 1) ? &tm_v : &tm_k, full(slot), h * HD, (item >> 1) * BKV, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint64_t qdesc = desc128(sQ);
      auto kw_of = [&](int t) { int r = Lk - t * BKV; r = r < BKV ? r : BKV; return (r + 31) & ~31; };
      auto issue_s = [&](int t) {
        const int item = 2 * t, slot = item % NSLOT;
        mbar_wait(full(slot), (uint32_t)(item / NSLOT) & 1u);
        tc_fence_after();
        const uint64_t kdesc = desc128(sRing + slot * TILE_BYTES);
        const uint32_t id = idesc(kw_of(t), 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma(tmem + S_COL, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), id, k != 0);
        tc_commit(empty(slot));
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(p_full, (uint32_t)t & 1u);               // P(t) in smem, S(t) consumed, O_t(t-1) folded
        tc_fence_after();
        // S(t+1) first: the softmax warps start tile t+1 while P V (t) is still executing
        if (t + 1 < nt) issue_s(t + 1);
        const int item = 2 * t + 1, slot = item % NSLOT;
        mbar_wait(full(slot), (uint32_t)(item / NSLOT) & 1u);
        tc_fence_after();
        const uint32_t vbase = sRing + slot * TILE_BYTES;
        const uint32_t id = idesc(HD, 1);
        const int ksteps = kw_of(t) / 16;
        for (int k = 0; k < ksteps; ++k) {
          // A = P: K-major, 64-key blocks of 16 KiB, 32 B per 16-key step inside a 128-byte row
          const uint64_t pdesc = desc128(sP + (uint32_t)((k >> 2) * TILE_BYTES + (k & 3) * 32));
          // B = V: MN-major, 16 keys = 16 rows of 128 B = 2048 B per step
          const uint64_t vdesc = desc128(vbase + (uint32_t)(k * 2048));
          tc_mma(tmem + O_COL, pdesc, vdesc, id, k != 0);
        }
        tc_commit(empty(slot));
        tc_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / output (8 warps, two threads per query row) =====================
    const int sw = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are visible to this warp
    const int hh = sw >> 2;                       // which 64-key half of the tile / 32-column half of the output
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + OFF_X);
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f, corr_prev = 0.f;
    auto fold = [&](int t_done, float c) {                 // o = o*c + O_t(t_done), O_t from the tensor core
      mbar_wait(o_full, (uint32_t)t_done & 1u);
      tc_fence_after();
      uint32_t raw[32];
      tc_ld32(lane_addr + O_COL + hh * 32, raw);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = fmaf(o[j], c, __uint_as_float(raw[j]));
      tc_fence_before();
    };
    for (int t = 0; t < nt; ++t) {
      const int valid = min(BKV, Lk - t * BKV) - hh * 64;      // valid keys in this thread's half (may be <= 0)
      const bool full_half = valid >= 64;
      mbar_wait(s_full, (uint32_t)t & 1u);
      tc_fence_after();
      // pass 1: max of the raw scores of this half (scale > 0 is applied afterwards)
      float mx = -INFINITY;
      if (full_half) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          tc_ld32(lane_addr + S_COL + hh * 64 + c * 32, raw);
          tc_wait_ld();
          float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;   // independent chains (ILP)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 = fmaxf(a0, __uint_as_float(raw[j]));
            a1 = fmaxf(a1, __uint_as_float(raw[j + 1]));
            a2 = fmaxf(a2, __uint_as_float(raw[j + 2]));
            a3 = fmaxf(a3, __uint_as_float(raw[j + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
        }
      } else {
        for (int c = 0; c * 32 < valid; ++c) {
          uint32_t raw[32];
          tc_ld32(lane_addr + S_COL + hh * 64 + c * 32, raw);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c * 32 + j < valid) ? __uint_as_float(raw[j]) : -INFINITY);
        }
      }
      xch[r * 2 + hh] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mt = fmaxf(m, fmaxf(xch[r * 2], xch[r * 2 + 1]) * scale_log2);
      const float corr = ex2(m - mt);                      // first tile: exp2(-inf) = 0
      m = mt;
      // P V (t-1) has long retired (it was issued before this tile's scores were read): fold it now; this
      // also guarantees the P buffer is free before pass 2 overwrites it
      if (t > 0) fold(t - 1, corr_prev);
      corr_prev = corr;
      // pass 2: p = exp2(s*scale - m), partial row sum, bf16 P into swizzled smem (row r, chunk j at j ^ (r & 7))
      float psum = 0.f;
      const uint32_t rowb = sP + (uint32_t)(hh * TILE_BYTES + r * 128);
      if (full_half) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          tc_ld32(lane_addr + S_COL + hh * 64 + c * 32, raw);
          tc_wait_ld();
          float p[32];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            p[j] = ex2(fmaf(__uint_as_float(raw[j]), scale_log2, -mt));
            p[j + 1] = ex2(fmaf(__uint_as_float(raw[j + 1]), scale_log2, -mt));
            p[j + 2] = ex2(fmaf(__uint_as_float(raw[j + 2]), scale_log2, -mt));
            p[j + 3] = ex2(fmaf(__uint_as_float(raw[j + 3]), scale_log2, -mt));
            s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3];
          }
          psum += (s0 + s1) + (s2 + s3);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(rowb + (uint32_t)((((c * 4 + j) ^ r) & 7) << 4), pack2(p[8 * j], p[8 * j + 1]), pack2(p[8 * j + 2], p[8 * j + 3]),
                   pack2(p[8 * j + 4], p[8 * j + 5]), pack2(p[8 * j + 6], p[8 * j + 7]));
        }
      } else {
        for (int c = 0; c * 32 < valid; ++c) {
          uint32_t raw[32];
          tc_ld32(lane_addr + S_COL + hh * 64 + c * 32, raw);
          tc_wait_ld();
          float p[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            p[j] = (c * 32 + j < valid) ? ex2(fmaf(__uint_as_float(raw[j]), scale_log2, -mt)) : 0.f;
            psum += p[j];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(rowb + (uint32_t)((((c * 4 + j) ^ r) & 7) << 4), pack2(p[8 * j], p[8 * j + 1]), pack2(p[8 * j + 2], p[8 * j + 3]),
                   pack2(p[8 * j + 4], p[8 * j + 5]), pack2(p[8 * j + 6], p[8 * j + 7]));
        }
      }
      l = l * corr + psum;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    fold(nt - 1, corr_prev);
    // ---- o / l -> bf16 -> swizzled staging tile (P block 0; the last P V product has retired) -> TMA store
    asm volatile("bar.sync 1, 256;" ::: "memory");         // everyone has read the last max exchange
    xch[r * 2 + hh] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (xch[r * 2] + xch[r * 2 + 1]);
    const uint32_t rowo = sP + (uint32_t)(r * 128);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(rowo + (uint32_t)((((hh * 4 + j) ^ r) & 7) << 4), pack2(o[8 * j] * inv, o[8 * j + 1] * inv),
             pack2(o[8 * j + 2] * inv, o[8 * j + 3] * inv), pack2(o[8 * j + 4] * inv, o[8 * j + 5] * inv),
             pack2(o[8 * j + 6] * inv, o[8 * j + 7] * inv));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight softmax warps only
    if (warp == 2 && lane == 0) {
      tma_store_3d(&tm_o, sP, h * HD, q0, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace fa

int attention_bf16_tc(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs, const void* V,
                      int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                      float scale, cudaStream_t st) {
  TCD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && qbs % 8 == 0 && kbs % 8 == 0 &&
              vbs % 8 == 0 && obs % 8 == 0, "tcd_attention(bf16): pitches and batch strides must be multiples of 8 elements");
  TCD_REQUIRE(((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0, "tcd_attention(bf16): 16-byte pointer alignment");
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap_3d_bf16(&tq, Q, (int64_t)heads * fa::HD, Lq, samples, ldq, qbs, fa::BQ))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, K, (int64_t)heads * fa::HD, Lk, samples, ldk, kbs, fa::BKV))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, V, (int64_t)heads * fa::HD, Lk, samples, ldv, vbs, fa::BKV))) return rc;
  if ((rc = make_tmap_3d_bf16(&to, O, (int64_t)heads * fa::HD, Lq, samples, ldo, obs, fa::BQ))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fa::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa::SMEM);
    if (e != cudaSuccess) { set_error("attention_tc: smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  dim3 grid(ceil_div(Lq, fa::BQ), heads, samples);
  fa::attention_tc_kernel<<<grid, fa::THREADS, fa::SMEM, st>>>(tq, tk, tv, to, Lq, Lk, scale * 1.4426950408889634f);
  return check_launch("attention_tc");
}

}  // namespace tcd
