// fp32 parity-mode kernels on the CUDA cores: C = act(A W^T + bias) and softmax(scale Q K^T) V.
// These carry the `TCD_F32` mode of tcd_gemm / tcd_attention, whose purpose is strict numerical
// parity with the reference's fp32 PyTorch evaluation (1e-4 relative, north_star); the
// performance path is the bf16 tcgen05 implementation in gemm_tc.cu / attention_tc.cu.
#include "common.cuh"

namespace tcd {

// ------------------------------------------------------------------------------------------
// 64x64x16 tile, 256 threads, 4x4 outputs per thread, both operands K-major (nn.Linear layout).
// ------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ W, int64_t ldw,
                                                       const float* __restrict__ bias, int act,
                                                       float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                       int64_t K) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Ws[GK][GN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  const int tr = tid / 16, tc = tid % 16;  // 16 x 16 thread grid, each 4 x 4
  const int lr = tid / 4, lk = (tid % 4) * 4;  // loader: row 0..63, k offset 0,4,8,12
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t k = k0 + lk + j;
      int64_t am = m0 + lr, wn = n0 + lr;
      As[lk + j][lr] = (am < M && k < K) ? __ldg(A + am * lda + k) : 0.f;
      Ws[lk + j][lr] = (wn < N && k < K) ? __ldg(W + wn * ldw + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][tr * 4]);
      float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tc * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + tr * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tc * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      C[m * ldc + n] = apply_act(v, act);
    }
  }
}

int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, int act, float* C,
             int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, act, C, ldc, M, N, K);
  return check_launch("gemm_f32");
}

// ------------------------------------------------------------------------------------------
// fp32 attention: one thread per query row (q and the output accumulator in registers), keys/values
// staged through shared memory 32 at a time, online softmax per key tile.  head_dim = 64.
// ------------------------------------------------------------------------------------------
constexpr int AQ = 128, AK = 32;

template <int HD>
__global__ void __launch_bounds__(AQ) attention_f32_kernel(
    const float* __restrict__ Q, int64_t ldq, int64_t qbs, const float* __restrict__ K, int64_t ldk, int64_t kbs,
    const float* __restrict__ V, int64_t ldv, int64_t vbs, float* __restrict__ O, int64_t ldo, int64_t obs,
    int heads, int Lq, int Lk, float scale) {
  __shared__ float Ks[AK][HD];
  __shared__ float Vs[AK][HD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * AQ + threadIdx.x;
  const bool active = qi < Lq;
  float q[HD], o[HD];
  const float* qp = Q + (int64_t)b * qbs + (int64_t)(active ? qi : 0) * ldq + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    float4 t = *reinterpret_cast<const float4*>(qp + d);
    q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
    o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += AK) {
    __syncthreads();
    for (int i = threadIdx.x; i < AK * HD / 4; i += AQ) {
      int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < Lk) {
        kv = *reinterpret_cast<const float4*>(K + (int64_t)b * kbs + (int64_t)(k0 + r) * ldk + h * HD + c);
        vv = *reinterpret_cast<const float4*>(V + (int64_t)b * vbs + (int64_t)(k0 + r) * ldv + h * HD + c);
      }
      *reinterpret_cast<float4*>(&Ks[r][c]) = kv;
      *reinterpret_cast<float4*>(&Vs[r][c]) = vv;
    }
    __syncthreads();
    float s[AK];
    float mt = mrun;
#pragma unroll
    for (int j = 0; j < AK; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) a = fmaf(q[d], Ks[j][d], a);
      s[j] = (k0 + j < Lk) ? a : -INFINITY;
      mt = fmaxf(mt, s[j]);
    }
    const float corr = expf(mrun - mt);  // exp(-inf) = 0 on the first tile
    lrun *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < AK; ++j) {
      float p = expf(s[j] - mt);
      lrun += p;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(p, Vs[j][d], o[d]);
    }
    mrun = mt;
  }
  if (active) {
    const float inv = 1.0f / lrun;
    float* op = O + (int64_t)b * obs + (int64_t)qi * ldo + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 4)
      *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
  }
}

int attention_f32(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs, const float* V,
                  int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                  float scale, cudaStream_t st) {
  dim3 grid(ceil_div(Lq, AQ), heads, samples);
  attention_f32_kernel<64><<<grid, AQ, 0, st>>>(Q, ldq, qbs, K, ldk, kbs, V, ldv, vbs, O, ldo, obs, heads, Lq, Lk, scale);
  return check_launch("attention_f32");
}

int attention_f32_hd32(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs, const float* V,
                       int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                       float scale, cudaStream_t st) {
  dim3 grid(ceil_div(Lq, AQ), heads, samples);
  attention_f32_kernel<32><<<grid, AQ, 0, st>>>(Q, ldq, qbs, K, ldk, kbs, V, ldv, vbs, O, ldo, obs, heads, Lq, Lk, scale);
  return check_launch("attention_f32_hd32");
}

}  // namespace tcd
