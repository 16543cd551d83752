// bf16 flash-style attention, head_dim 64, no mask: O = softmax(scale * Q K^T) V per (sample, head).
// Round-1 tensor-core path: mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix-fed fragments,
// cp.async double-buffered K/V tiles in XOR-swizzled shared memory, online softmax in fp32 registers,
// scores never touch HBM (the reference materialises (B,8,L,L) fp32, model/model.py:97-102).
// The tcgen05/TMEM version of this kernel is the planned replacement.
//
// CTA = 4 warps = 64 query rows of one (sample, head); K/V streamed 64 keys at a time.
#include "common.cuh"

namespace tcd {

constexpr int FQ = 64, FK = 64, FD = 64;
constexpr int FA_THREADS = 128;

__device__ __forceinline__ uint32_t sw_off(int r, int c) {  // byte offset of element (r, c) in a [rows][64] bf16 tile
  return (uint32_t)(r * 128 + ((((c >> 3) ^ r) & 7) << 4) + ((c & 7) << 1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// cooperative load of a [64][64] bf16 tile (rows row0.., of a matrix with `nrows` valid rows) into swizzled smem
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const __nv_bfloat16* __restrict__ g, int64_t ld,
                                          int row0, int nrows) {
#pragma unroll
  for (int i = 0; i < (FQ * 8) / FA_THREADS; ++i) {
    int idx = threadIdx.x + i * FA_THREADS;
    int r = idx >> 3, ch = idx & 7;
    bool ok = row0 + r < nrows;
    const __nv_bfloat16* src = g + (int64_t)(ok ? row0 + r : 0) * ld + ch * 8;
    cp_async16(smem_tile + (uint32_t)(r * 128 + (((ch ^ r) & 7) << 4)), src, ok);
  }
}

__global__ void __launch_bounds__(FA_THREADS) attention_bf16_kernel(
    const __nv_bfloat16* __restrict__ Q, int64_t ldq, int64_t qbs, const __nv_bfloat16* __restrict__ K, int64_t ldk,
    int64_t kbs, const __nv_bfloat16* __restrict__ V, int64_t ldv, int64_t vbs, __nv_bfloat16* __restrict__ O,
    int64_t ldo, int64_t obs, int Lq, int Lk, float scale_log2) {
  __shared__ __align__(1024) uint8_t smem[FQ * 128 + 2 * 2 * FK * 128];  // Q | K0 V0 | K1 V1  (40 KiB)
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sKV = sQ + FQ * 128;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qg = Q + (int64_t)b * qbs + h * FD;
  const __nv_bfloat16* kg = K + (int64_t)b * kbs + h * FD;
  const __nv_bfloat16* vg = V + (int64_t)b * vbs + h * FD;

  load_tile(sQ, qg, ldq, q0, Lq);
  load_tile(sKV, kg, ldk, 0, Lk);
  load_tile(sKV + FK * 128, vg, ldv, 0, Lk);
  cp_async_commit();

  const int ntiles = (Lk + FK - 1) / FK;
  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  for (int t = 0; t < ntiles; ++t) {
    const uint32_t sK = sKV + (uint32_t)((t & 1) * 2 * FK * 128), sV = sK + FK * 128;
    if (t + 1 < ntiles) {  // prefetch next K/V tile into the other buffer
      const uint32_t nK = sKV + (uint32_t)(((t + 1) & 1) * 2 * FK * 128);
      load_tile(nK, kg, ldk, (t + 1) * FK, Lk);
      load_tile(nK + FK * 128, vg, ldv, (t + 1) * FK, Lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldsm_x4(sQ + sw_off(warp * 16 + (lane & 15), kk * 16 + ((lane >> 4) << 3)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int key = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int d = kk * 16 + (((lane >> 3) & 1) << 3);
        ldsm_x4(sK + sw_off(key, d), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[kk], b0, b1);
        mma_bf16(s[2 * np + 1], qf[kk], b2, b3);
      }
    }
    // ---- mask keys beyond Lk, online softmax (rows lane/4 and lane/4 + 8) ----
    const int kbase = t * FK + 2 * (lane & 3);
    float mt[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + nt * 8 + (e & 1);
        float v = key < Lk ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        mt[e >> 1] = fmaxf(mt[e >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 1));
      mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 2));
    }
    float corr[2], psum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f(mrow[r] - mt[r]);
      mrow[r] = mt[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float p0 = exp2f(s[nt][0] - mt[0]), p1 = exp2f(s[nt][1] - mt[0]);
      float p2 = exp2f(s[nt][2] - mt[1]), p3 = exp2f(s[nt][3] - mt[1]);
      psum[0] += p0 + p1;
      psum[1] += p2 + p3;
      // accumulator tiles (2j, 2j+1) of S form the A fragment of key block j for the P V product
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * corr[r] + psum[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        const int key = kt * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int d = dp * 16 + ((lane >> 4) << 3);
        ldsm_x4_t(sV + sw_off(key, d), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[kt], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[kt], b2, b3);
      }
    }
    __syncthreads();  // all warps done with this K/V buffer before it is refilled
  }
  // ---- finalize: divide by the row sums (quad-reduced) and store bf16 ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* og = O + (int64_t)b * obs + h * FD + 2 * (lane & 3);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (r0 < Lq) *reinterpret_cast<uint32_t*>(og + (int64_t)r0 * ldo + nt * 8) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
    if (r1 < Lq) *reinterpret_cast<uint32_t*>(og + (int64_t)r1 * ldo + nt * 8) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

int attention_bf16(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs, const void* V,
                   int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                   float scale, cudaStream_t st) {
  TCD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "tcd_attention(bf16): row pitches must be multiples of 8 elements");
  TCD_REQUIRE(((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V) % 16 == 0 && (uintptr_t)O % 4 == 0, "tcd_attention(bf16): pointer alignment");
  TCD_REQUIRE(qbs % 8 == 0 && kbs % 8 == 0 && vbs % 8 == 0 && obs % 2 == 0, "tcd_attention(bf16): batch strides must be multiples of 8 elements");
  dim3 grid(ceil_div(Lq, FQ), heads, samples);
  attention_bf16_kernel<<<grid, FA_THREADS, 0, st>>>(
      (const __nv_bfloat16*)Q, ldq, qbs, (const __nv_bfloat16*)K, ldk, kbs, (const __nv_bfloat16*)V, ldv, vbs,
      (__nv_bfloat16*)O, ldo, obs, Lq, Lk, scale * 1.4426950408889634f);
  return check_launch("attention_bf16");
}

}  // namespace tcd
