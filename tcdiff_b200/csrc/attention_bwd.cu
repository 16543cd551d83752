// Attention backward (fp32, CUDA cores): dQ, dK, dV of O = softmax(scale Q K^T) V per (sample, head), head_dim 64.
// First correct version of the training path's attention gradient (flash-style: P is recomputed from the saved
// log-sum-exp, nothing of size Lq x Lk touches HBM); the tensor-core version is future work.
//   stats : lse_i = log sum_j exp(scale q_i.k_j),  delta_i = dO_i . O_i
//   dkdv  : block = 32 keys, thread = (key, 16-dim quarter); loops over all queries
//           p = exp(s - lse); dV_j += p dO_i; dP = dO_i.V_j; dS = p (dP - delta_i); dK_j += scale dS q_i
//   dq    : block = 32 queries, thread = (query, quarter); loops over all keys; dQ_i += scale dS k_j
// Row layouts are the forward's (pitch + head*64 column offset + batch stride), so gradients can be written straight
// into packed [Q|K] gradient buffers.
#include "common.cuh"

namespace tcd {

constexpr int BD = 64;

struct AttnPtr {
  const float* p; int64_t ld, bs;
};
struct AttnOut {
  float* p; int64_t ld, bs;
};

__global__ void __launch_bounds__(128) attn_bwd_stats_kernel(AttnPtr Q, AttnPtr K, AttnPtr O, AttnPtr dO, float* __restrict__ lse,
                                                             float* __restrict__ delta, int Lq, int Lk, float scale) {
  __shared__ float Ks[32][BD];
  const int b = blockIdx.z, h = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
  const bool active = i < Lq;
  float q[BD];
  const float* qp = Q.p + (int64_t)b * Q.bs + (int64_t)(active ? i : 0) * Q.ld + h * BD;
#pragma unroll
  for (int d = 0; d < BD; ++d) q[d] = qp[d] * scale;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < Lk; k0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * BD; e += 128) {
      const int r = e / BD, c = e % BD;
      Ks[r][c] = (k0 + r < Lk) ? K.p[(int64_t)b * K.bs + (int64_t)(k0 + r) * K.ld + h * BD + c] : 0.f;
    }
    __syncthreads();
    float s[32], mt = m;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < BD; ++d) a = fmaf(q[d], Ks[j][d], a);
      s[j] = (k0 + j < Lk) ? a : -INFINITY;
      mt = fmaxf(mt, s[j]);
    }
    l *= expf(m - mt);
#pragma unroll
    for (int j = 0; j < 32; ++j) l += expf(s[j] - mt);
    m = mt;
  }
  if (active) {
    const float* op = O.p + (int64_t)b * O.bs + (int64_t)i * O.ld + h * BD;
    const float* gp = dO.p + (int64_t)b * dO.bs + (int64_t)i * dO.ld + h * BD;
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < BD; ++d) dl = fmaf(op[d], gp[d], dl);
    const int64_t o = ((int64_t)b * gridDim.y + h) * Lq + i;
    lse[o] = m + logf(l);
    delta[o] = dl;
  }
}

__device__ __forceinline__ float quad_sum(float v) {   // the 4 dim-quarters of a row sit in adjacent lanes
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

__global__ void __launch_bounds__(128) attn_bwd_dkdv_kernel(AttnPtr Q, AttnPtr K, AttnPtr V, AttnPtr dO, const float* __restrict__ lse,
                                                            const float* __restrict__ delta, AttnOut dK, AttnOut dV, int Lq, int Lk,
                                                            float scale) {
  __shared__ float Qs[32][BD], Gs[32][BD], Ls[32], Ds[32];
  const int b = blockIdx.z, h = blockIdx.y;
  const int jl = threadIdx.x >> 2, qd = threadIdx.x & 3;       // key within the tile, 16-dim quarter
  const int j = blockIdx.x * 32 + jl;
  const bool active = j < Lk;
  float k[16], v[16], dk[16], dv[16];
  const int64_t koff = (int64_t)b * K.bs + (int64_t)(active ? j : 0) * K.ld + h * BD + qd * 16;
  const int64_t voff = (int64_t)b * V.bs + (int64_t)(active ? j : 0) * V.ld + h * BD + qd * 16;
#pragma unroll
  for (int d = 0; d < 16; ++d) { k[d] = K.p[koff + d]; v[d] = V.p[voff + d]; dk[d] = 0.f; dv[d] = 0.f; }
  const int64_t sbase = ((int64_t)b * gridDim.y + h) * Lq;
  for (int i0 = 0; i0 < Lq; i0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * BD; e += 128) {
      const int r = e / BD, c = e % BD;
      const bool ok = i0 + r < Lq;
      Qs[r][c] = ok ? Q.p[(int64_t)b * Q.bs + (int64_t)(i0 + r) * Q.ld + h * BD + c] : 0.f;
      Gs[r][c] = ok ? dO.p[(int64_t)b * dO.bs + (int64_t)(i0 + r) * dO.ld + h * BD + c] : 0.f;
    }
    if (threadIdx.x < 32) {
      const bool ok = i0 + threadIdx.x < Lq;
      Ls[threadIdx.x] = ok ? lse[sbase + i0 + threadIdx.x] : INFINITY;   // p = exp(s - inf) = 0 for absent queries
      Ds[threadIdx.x] = ok ? delta[sbase + i0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        s = fmaf(Qs[i][qd * 16 + d], k[d], s);
        dp = fmaf(Gs[i][qd * 16 + d], v[d], dp);
      }
      s = quad_sum(s) * scale;
      dp = quad_sum(dp);
      const float p = expf(s - Ls[i]);
      const float ds = p * (dp - Ds[i]) * scale;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        dv[d] = fmaf(p, Gs[i][qd * 16 + d], dv[d]);
        dk[d] = fmaf(ds, Qs[i][qd * 16 + d], dk[d]);
      }
    }
  }
  if (active) {
    float* okp = dK.p + (int64_t)b * dK.bs + (int64_t)j * dK.ld + h * BD + qd * 16;
    float* ovp = dV.p + (int64_t)b * dV.bs + (int64_t)j * dV.ld + h * BD + qd * 16;
#pragma unroll
    for (int d = 0; d < 16; ++d) { okp[d] = dk[d]; ovp[d] = dv[d]; }
  }
}

__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(AttnPtr Q, AttnPtr K, AttnPtr V, AttnPtr dO, const float* __restrict__ lse,
                                                          const float* __restrict__ delta, AttnOut dQ, int Lq, int Lk, float scale) {
  __shared__ float Ks[32][BD], Vs[32][BD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int il = threadIdx.x >> 2, qd = threadIdx.x & 3;
  const int i = blockIdx.x * 32 + il;
  const bool active = i < Lq;
  float q[16], g[16], dq[16];
  const int64_t qoff = (int64_t)b * Q.bs + (int64_t)(active ? i : 0) * Q.ld + h * BD + qd * 16;
  const int64_t goff = (int64_t)b * dO.bs + (int64_t)(active ? i : 0) * dO.ld + h * BD + qd * 16;
#pragma unroll
  for (int d = 0; d < 16; ++d) { q[d] = Q.p[qoff + d]; g[d] = dO.p[goff + d]; dq[d] = 0.f; }
  const int64_t sidx = ((int64_t)b * gridDim.y + h) * Lq + (active ? i : 0);
  const float li = lse[sidx], di = delta[sidx];
  for (int k0 = 0; k0 < Lk; k0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * BD; e += 128) {
      const int r = e / BD, c = e % BD;
      const bool ok = k0 + r < Lk;
      Ks[r][c] = ok ? K.p[(int64_t)b * K.bs + (int64_t)(k0 + r) * K.ld + h * BD + c] : 0.f;
      Vs[r][c] = ok ? V.p[(int64_t)b * V.bs + (int64_t)(k0 + r) * V.ld + h * BD + c] : 0.f;
    }
    __syncthreads();
    const int nk = min(32, Lk - k0);
    for (int j = 0; j < nk; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        s = fmaf(q[d], Ks[j][qd * 16 + d], s);
        dp = fmaf(g[d], Vs[j][qd * 16 + d], dp);
      }
      s = quad_sum(s) * scale;
      dp = quad_sum(dp);
      const float ds = expf(s - li) * (dp - di) * scale;
#pragma unroll
      for (int d = 0; d < 16; ++d) dq[d] = fmaf(ds, Ks[j][qd * 16 + d], dq[d]);
    }
  }
  if (active) {
    float* op = dQ.p + (int64_t)b * dQ.bs + (int64_t)i * dQ.ld + h * BD + qd * 16;
#pragma unroll
    for (int d = 0; d < 16; ++d) op[d] = dq[d];
  }
}

}  // namespace tcd

using namespace tcd;

extern "C" int64_t tcd_attention_backward_workspace_floats(int samples, int heads, int Lq) {
  return 2LL * samples * heads * Lq;
}

extern "C" int tcd_attention_backward(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs,
                                      const float* V, int64_t ldv, int64_t vbs, const float* O, int64_t ldo, int64_t obs,
                                      const float* dO, int64_t ldg, int64_t gbs, float* dQ, int64_t lddq, int64_t dqbs,
                                      float* dK, int64_t lddk, int64_t dkbs, float* dV, int64_t lddv, int64_t dvbs,
                                      float* workspace, int samples, int heads, int Lq, int Lk, float scale, void* stream) {
  TCD_REQUIRE(samples >= 0 && heads > 0 && Lq >= 0 && Lk > 0, "tcd_attention_backward: bad shape");
  if (samples == 0 || Lq == 0) return TCD_OK;
  TCD_REQUIRE(Q && K && V && O && dO && dQ && dK && dV && workspace, "tcd_attention_backward: null pointer");
  TCD_REQUIRE(heads <= 65535 && samples <= 65535, "tcd_attention_backward: grid limit");
  cudaStream_t st = as_stream(stream);
  float* lse = workspace;
  float* delta = workspace + (int64_t)samples * heads * Lq;
  AttnPtr q{Q, ldq, qbs}, k{K, ldk, kbs}, v{V, ldv, vbs}, o{O, ldo, obs}, g{dO, ldg, gbs};
  attn_bwd_stats_kernel<<<dim3(ceil_div(Lq, 128), heads, samples), 128, 0, st>>>(q, k, o, g, lse, delta, Lq, Lk, scale);
  int rc = check_launch("attn_bwd_stats");
  if (rc) return rc;
  attn_bwd_dkdv_kernel<<<dim3(ceil_div(Lk, 32), heads, samples), 128, 0, st>>>(q, k, v, g, lse, delta, AttnOut{dK, lddk, dkbs},
                                                                              AttnOut{dV, lddv, dvbs}, Lq, Lk, scale);
  rc = check_launch("attn_bwd_dkdv");
  if (rc) return rc;
  attn_bwd_dq_kernel<<<dim3(ceil_div(Lq, 32), heads, samples), 128, 0, st>>>(q, k, v, g, lse, delta, AttnOut{dQ, lddq, dqbs}, Lq, Lk,
                                                                            scale);
  return check_launch("attn_bwd_dq");
}
