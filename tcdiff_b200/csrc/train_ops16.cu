// Backward-pass primitives of the bf16 training tape: GEMM operands and their gradients live in bf16 (so that every
// dgrad / wgrad contraction reads them as stored, no casts), the residual stream and all parameter gradients stay fp32.
//   tcd_act_forward_bf16 / tcd_act_backward_bf16   ReLU / GELU(erf) / Mish / SiLU on bf16 pre-activations
//   tcd_layernorm_backward_mixed  LayerNorm backward with bf16 upstream gradients, optionally with the rotary branch
//                                 fused (dy_eff = dy + R(-theta) dy_rot), dx in fp32 (residual stream) or bf16
//   tcd_film_backward_bf16        featurewise-affine backward with bf16 v / dv, row-chunk parallel + deterministic reduce
//   tcd_colsum_bf16               bias gradients: column sums of a bf16 (tokens x features) gradient
// fp32 counterparts: train_ops.cu.  Reference ops: model/model.py:171-173 (FiLM), nn.LayerNorm, F.gelu, nn.Mish,
// nn.SiLU, model/rotary_embedding_torch.py:39-59.
#include "common.cuh"
#include "dropout.cuh"

namespace tcd {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// optional fused dropout of a kernel's bf16 operand: mask of (state, site) indexed by the operand's flat element index
struct DropArg {
  const uint64_t* state;     // nullptr = no dropout
  uint32_t site, thr;
  float rk;
};
__device__ __forceinline__ float drop_scale(const DropArg& d, uint32_t seed, int64_t i) {
  return drop_keep(seed, (uint32_t)i, (uint32_t)(i >> 32), d.thr) ? d.rk : 0.f;
}
static DropArg make_drop(float p, const void* rng_state, uint32_t site) {
  DropArg d;
  d.state = p > 0.f ? (const uint64_t*)rng_state : nullptr;
  d.site = site;
  d.thr = drop_threshold(p);
  d.rk = 1.0f / (1.0f - p);
  return d;
}

// ---------------------------------------------------------------------------------------- activations (bf16)
__device__ __forceinline__ float act_grad16(float z, int act) {
  switch (act) {
    case TCD_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case TCD_ACT_GELU: {
      const float phi = 0.3989422804014327f * __expf(-0.5f * z * z);
      return 0.5f * (1.0f + erff(z * 0.70710678118654752440f)) + z * phi;
    }
    case TCD_ACT_MISH: {
      const float sp = softplus_t(z), th = tanhf(sp);
      const float sg = 1.0f / (1.0f + __expf(-z));
      return th + z * (1.0f - th * th) * sg;
    }
    case TCD_ACT_SILU: {
      const float sg = 1.0f / (1.0f + __expf(-z));
      return sg * (1.0f + z * (1.0f - sg));
    }
    default: return 1.f;
  }
}

// GELU (exact-erf form) and its derivative for bf16 tensors, sharing one rcp.approx and one ex2.approx per element:
// Phi(-|z|) = 0.5 erfc(|z|/sqrt2) = 0.5 poly(t) exp(-z^2/2), t = 1/(1 + p |z|/sqrt2) (Abramowitz-Stegun 7.1.26, |err| < 1.5e-7,
// the form the sampler's GEMM epilogue uses), and the same exp(-z^2/2) is the density of the derivative Phi(z) + z phi(z).
// erff() + __expf() made the forward launch of the c3 step (96 000 x 1024) 192 us against an HBM floor of 60 us (r01 profile).
__device__ __forceinline__ void gelu_parts16(float z, float& cdf, float& e) {
  const float a = fabsf(z) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
  const float p = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * a * a));      // exp(-z^2 / 2)
  const float h = 0.5f * p * e;                                                           // Phi(-|z|)
  cdf = z >= 0.f ? 1.0f - h : h;
}

// forward: out = act(z) * m; backward: out = dy * m * act'(z)  (m = fused dropout mask / (1-p) of the activation's OUTPUT)
// GELU: the hot instantiation (linear1 of every layer) with the activation as a compile-time constant
template <bool BWD, bool GELU>
__global__ void __launch_bounds__(256) act16_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy,
                                                    __nv_bfloat16* __restrict__ out, int64_t n, int act, DropArg dr) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const uint32_t dseed = dr.state ? drop_site_seed(dr.state, dr.site) : 0u;
  auto fwd = [&](float v) -> float {
    if constexpr (GELU) { float c, e; gelu_parts16(v, c, e); return v * c; }
    else return apply_act(v, act);
  };
  auto grad = [&](float v) -> float {
    if constexpr (GELU) { float c, e; gelu_parts16(v, c, e); return fmaf(v * 0.3989422804014327f, e, c); }
    else return act_grad16(v, act);
  };
  if (i + 8 <= n) {
    const uint4 zu = *reinterpret_cast<const uint4*>(z + i);
    uint4 gu = make_uint4(0, 0, 0, 0);
    if (BWD) gu = *reinterpret_cast<const uint4*>(dy + i);
    const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&zu);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gu);
    uint4 ou;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ou);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 zf = __bfloat1622float2(z2[j]);
      float2 r;
      if (BWD) {
        const float2 gf = __bfloat1622float2(g2[j]);
        r = make_float2(gf.x * grad(zf.x), gf.y * grad(zf.y));
      } else {
        r = make_float2(fwd(zf.x), fwd(zf.y));
      }
      if (dr.state) { r.x *= drop_scale(dr, dseed, i + 2 * j); r.y *= drop_scale(dr, dseed, i + 2 * j + 1); }
      o2[j] = __floats2bfloat162_rn(r.x, r.y);
    }
    *reinterpret_cast<uint4*>(out + i) = ou;
  } else {
    for (int64_t k = i; k < n; ++k) {
      const float zf = __bfloat162float(z[k]);
      const float m = dr.state ? drop_scale(dr, dseed, k) : 1.0f;
      out[k] = __float2bfloat16_rn((BWD ? __bfloat162float(dy[k]) * grad(zf) : fwd(zf)) * m);
    }
  }
}

// ---------------------------------------------------------------------------------------- elementwise dropout
// y = x * keep / (1 - p); the same launch on dy is the backward pass.  8 bf16 (or 4 fp32) elements per thread.
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, uint32_t thr,
                                                      float rk, const uint64_t* __restrict__ state, uint32_t site) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint32_t seed = drop_site_seed(state, site);
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i0 >= n) return;
  float v[V];
  if (i0 + V <= n) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i0);
    if constexpr (sizeof(T) == 2) {
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(p2[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    } else {
      v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int64_t i = i0 + j;
      v[j] = drop_keep(seed, (uint32_t)i, (uint32_t)(i >> 32), thr) ? v[j] * rk : 0.f;
    }
    uint4 o;
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    } else {
      o = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
    *reinterpret_cast<uint4*>(y + i0) = o;
  } else {
    for (int64_t i = i0; i < n; ++i) {
      const float f = Conv<T>::from(x[i]);
      y[i] = Conv<T>::to(drop_keep(seed, (uint32_t)i, (uint32_t)(i >> 32), thr) ? f * rk : 0.f);
    }
  }
}

// test/debug: the attention-probability mask (0 or 1/(1-p)) as fp32 (samples, heads, Lq, Lk)
__global__ void __launch_bounds__(256) dropout_mask_attention_kernel(float* __restrict__ out, int64_t rows, int Lk, uint32_t thr,
                                                                     float rk, const uint64_t* __restrict__ state, uint32_t site) {
  const uint32_t seed = drop_site_seed(state, site);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Lk) return;
  const uint32_t row = (uint32_t)(i / Lk), k = (uint32_t)(i % Lk);
  out[i] = drop_keep(seed, row, k, thr) ? rk : 0.f;
}

// ---------------------------------------------------------------------------------------- LayerNorm backward (mixed)
// warp per row (grid-stride); lane owns columns {4*lane + 128*k .. +3}.  dy (and the optional gradient of the rotated
// copy, rotated back by -theta) in DYT; x, dx and the optional residual-path gradient dres (added to dx) in XT;
// per-BLOCK partial dgamma/dbeta rows in fp32 (tcd_layernorm_backward_mixed_partials(rows) of them).
template <typename XT, typename DYT, int NV>
__global__ void __launch_bounds__(256) layernorm_backward_mixed_kernel(
    const XT* __restrict__ x, const float* __restrict__ gamma, const DYT* __restrict__ dy, const DYT* __restrict__ dyrot,
    const float* __restrict__ rot_cos, const float* __restrict__ rot_sin, int tps, float eps, const XT* __restrict__ dres,
    XT* __restrict__ dx, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, int64_t rows, DropArg dr) {
  constexpr int D = 128 * NV;
  constexpr float invD = 1.0f / D;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = dr.state ? drop_site_seed(dr.state, dr.site) : 0u;      // dropout fused on the norm's INPUT
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp; row < rows; row += nwarps) {
    float4 xv[NV], dv[NV], mk[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int64_t e0 = row * D + 4 * (lane + 32 * k);
      xv[k] = ld4(x + e0);
      if (dr.state) {
        mk[k] = make_float4(drop_scale(dr, dseed, e0), drop_scale(dr, dseed, e0 + 1), drop_scale(dr, dseed, e0 + 2),
                            drop_scale(dr, dseed, e0 + 3));
        xv[k].x *= mk[k].x; xv[k].y *= mk[k].y; xv[k].z *= mk[k].z; xv[k].w *= mk[k].w;
      }
      dv[k] = dy ? ld4(dy + e0) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (xv[k].x + xv[k].y) + (xv[k].z + xv[k].w);
    }
    if (dyrot) {
      const int pos = (int)(row % tps);
      const float* cr = rot_cos + (int64_t)pos * (D / 2);
      const float* sr = rot_sin + (int64_t)pos * (D / 2);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const float4 r = ld4(dyrot + row * D + 4 * (lane + 32 * k));
        const float2 c = __ldg(reinterpret_cast<const float2*>(cr) + lane + 32 * k);
        const float2 sn = __ldg(reinterpret_cast<const float2*>(sr) + lane + 32 * k);
        dv[k].x += r.x * c.x + r.y * sn.x;       // transpose of (x c - y s, y c + x s)
        dv[k].y += r.y * c.x - r.x * sn.x;
        dv[k].z += r.z * c.y + r.w * sn.y;
        dv[k].w += r.w * c.y - r.z * sn.y;
      }
    }
    const float mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      xv[k].x -= mean; xv[k].y -= mean; xv[k].z -= mean; xv[k].w -= mean;
      q += (xv[k].x * xv[k].x + xv[k].y * xv[k].y) + (xv[k].z * xv[k].z + xv[k].w * xv[k].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invD + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      xv[k].x *= rstd; xv[k].y *= rstd; xv[k].z *= rstd; xv[k].w *= rstd;
      ag[k].x += dv[k].x * xv[k].x; ag[k].y += dv[k].y * xv[k].y; ag[k].z += dv[k].z * xv[k].z; ag[k].w += dv[k].w * xv[k].w;
      ab[k].x += dv[k].x; ab[k].y += dv[k].y; ab[k].z += dv[k].z; ab[k].w += dv[k].w;
      dv[k].x *= g[k].x; dv[k].y *= g[k].y; dv[k].z *= g[k].z; dv[k].w *= g[k].w;
      s1 += (dv[k].x + dv[k].y) + (dv[k].z + dv[k].w);
      s2 += (dv[k].x * xv[k].x + dv[k].y * xv[k].y) + (dv[k].z * xv[k].z + dv[k].w * xv[k].w);
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float4 o;
      o.x = rstd * (dv[k].x - s1 - xv[k].x * s2); o.y = rstd * (dv[k].y - s1 - xv[k].y * s2);
      o.z = rstd * (dv[k].z - s1 - xv[k].z * s2); o.w = rstd * (dv[k].w - s1 - xv[k].w * s2);
      if (dres) {                                  // gradient arriving through the residual connection around the norm
        const float4 e = ld4(dres + row * D + 4 * (lane + 32 * k));
        o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
      }
      if (dr.state) { o.x *= mk[k].x; o.y *= mk[k].y; o.z *= mk[k].z; o.w *= mk[k].w; }
      st4(dx + row * D + 4 * (lane + 32 * k), o);
    }
  }
  // block-level reduction of the 8 warps' partial dgamma / dbeta rows -> ONE partial row per block
  __shared__ float red[8][D];
  const int wib = threadIdx.x >> 5;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int k = 0; k < NV; ++k) reinterpret_cast<float4*>(red[wib])[lane + 32 * k] = pass == 0 ? ag[k] : ab[k];
    __syncthreads();
    float* dst = (pass == 0 ? dgamma_part : dbeta_part) + (int64_t)blockIdx.x * D;
    for (int c = threadIdx.x; c < D; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][c];
      dst[c] = t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------- LayerNorm forward, bf16 in/out
// (the per-head-merge LayerNorm(eps=1e-6) of SBI_MSA applied to the bf16 output of the fc projection, model.py:104-107)
template <int NV>
__global__ void __launch_bounds__(256) layernorm16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          __nv_bfloat16* __restrict__ y, int64_t rows, DropArg dr) {
  constexpr int D = 128 * NV;
  constexpr float invD = 1.0f / D;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint32_t dseed = dr.state ? drop_site_seed(dr.state, dr.site) : 0u;      // dropout fused on the norm's INPUT
  float4 xv[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int64_t e0 = row * D + 4 * (lane + 32 * k);
    xv[k] = ld4(x + e0);
    if (dr.state) {
      xv[k].x *= drop_scale(dr, dseed, e0); xv[k].y *= drop_scale(dr, dseed, e0 + 1);
      xv[k].z *= drop_scale(dr, dseed, e0 + 2); xv[k].w *= drop_scale(dr, dseed, e0 + 3);
    }
    s += (xv[k].x + xv[k].y) + (xv[k].z + xv[k].w);
  }
  const float mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    xv[k].x -= mean; xv[k].y -= mean; xv[k].z -= mean; xv[k].w -= mean;
    q += (xv[k].x * xv[k].x + xv[k].y * xv[k].y) + (xv[k].z * xv[k].z + xv[k].w * xv[k].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * invD + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * k);
    st4(y + row * D + 4 * (lane + 32 * k), make_float4(xv[k].x * rstd * g.x + b.x, xv[k].y * rstd * g.y + b.y,
                                                      xv[k].z * rstd * g.z + b.z, xv[k].w * rstd * g.w + b.w));
  }
}

// ---------------------------------------------------------------------------------------- FiLM backward (bf16 v)
// out = x + (1 + scale[b]) v + shift[b].  Block = (row chunk, sample), thread = 4 columns; partial (dscale | dshift)
// rows per chunk, reduced afterwards in a fixed order.
constexpr int kFilmChunk = 25;
__global__ void __launch_bounds__(256) film_backward16_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ v,
                                                              const float* __restrict__ film, int64_t film_ld, int64_t film_off,
                                                              __nv_bfloat16* __restrict__ dv, float* __restrict__ part, int L, int D,
                                                              int chunks, DropArg dr) {
  const int b = blockIdx.y, ch = blockIdx.x, c = threadIdx.x * 4;
  if (c >= D) return;
  const uint32_t dseed = dr.state ? drop_site_seed(dr.state, dr.site) : 0u;      // dropout fused on v
  const float4 sc4 = ld4(film + b * film_ld + film_off + c);
  const float4 sc = make_float4(1.f + sc4.x, 1.f + sc4.y, 1.f + sc4.z, 1.f + sc4.w);
  float4 ds = make_float4(0.f, 0.f, 0.f, 0.f), dsh = ds;
  const int r0 = ch * kFilmChunk, r1 = min(L, r0 + kFilmChunk);
  for (int r = r0; r < r1; ++r) {
    const int64_t o = ((int64_t)b * L + r) * D + c;
    const float4 g = ld4(dout + o);
    float4 vv = ld4(v + o);
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (dr.state) {
      m = make_float4(drop_scale(dr, dseed, o), drop_scale(dr, dseed, o + 1), drop_scale(dr, dseed, o + 2), drop_scale(dr, dseed, o + 3));
      vv.x *= m.x; vv.y *= m.y; vv.z *= m.z; vv.w *= m.w;
    }
    st4(dv + o, make_float4(sc.x * g.x * m.x, sc.y * g.y * m.y, sc.z * g.z * m.z, sc.w * g.w * m.w));
    ds.x += g.x * vv.x; ds.y += g.y * vv.y; ds.z += g.z * vv.z; ds.w += g.w * vv.w;
    dsh.x += g.x; dsh.y += g.y; dsh.z += g.z; dsh.w += g.w;
  }
  float* p = part + ((int64_t)b * chunks + ch) * 2 * D;
  st4(p + c, ds);
  st4(p + D + c, dsh);
}
// forward of the same block for the training tape: out = x + (1 + scale[b]) (v m) + shift[b]  (film == nullptr: x + v m)
__global__ void __launch_bounds__(256) film_residual16_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                                              const float* __restrict__ film, int64_t film_ld, int64_t film_off,
                                                              float* __restrict__ out, int64_t rows, int L, int D, DropArg dr) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // float4 index
  const int dq = D / 4;
  if (q >= rows * dq) return;
  const int64_t row = q / dq;
  const int c = (int)(q % dq) * 4;
  const int64_t o = row * D + c;
  float4 vv = ld4(v + o);
  if (dr.state) {
    const uint32_t dseed = drop_site_seed(dr.state, dr.site);
    vv.x *= drop_scale(dr, dseed, o); vv.y *= drop_scale(dr, dseed, o + 1);
    vv.z *= drop_scale(dr, dseed, o + 2); vv.w *= drop_scale(dr, dseed, o + 3);
  }
  const float4 xv = ld4(x + o);
  float4 r;
  if (film) {
    const float* f = film + (row / L) * film_ld + film_off + c;
    const float4 sc = ld4(f), sh = ld4(f + D);
    r = make_float4(xv.x + (1.f + sc.x) * vv.x + sh.x, xv.y + (1.f + sc.y) * vv.y + sh.y, xv.z + (1.f + sc.z) * vv.z + sh.z,
                    xv.w + (1.f + sc.w) * vv.w + sh.w);
  } else {
    r = make_float4(xv.x + vv.x, xv.y + vv.y, xv.z + vv.z, xv.w + vv.w);
  }
  st4(out + o, r);
}

__global__ void __launch_bounds__(256) film_reduce_kernel(const float* __restrict__ part, float* __restrict__ dfilm,
                                                          int64_t dfilm_ld, int64_t dfilm_off, int chunks, int D2) {
  const int b = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c >= D2) return;
  const float* p = part + (int64_t)b * chunks * D2 + c;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += p[(int64_t)k * D2];
  dfilm[b * dfilm_ld + dfilm_off + c] = acc;
}

// ---------------------------------------------------------------------------------------- column sums of a bf16 matrix
// block = (256 columns, 512-row chunk): thread = 8 columns (one 16-byte load) x one of 8 row lanes, four rows in flight per
// thread; partial rows reduced by a second pass.  (r01: 4-byte loads, one in flight per thread: 41 us per bias gradient of
// the c3 step against an HBM floor of 15-30 us, 2.2 ms per step in 53 launches.)  Pitches that are not a multiple of 8
// elements, or a base that is not 16-byte aligned, take the narrow path.
constexpr int kColsumRows = 512;
template <bool WIDE>
__global__ void __launch_bounds__(256) colsum16_partial_kernel(const __nv_bfloat16* __restrict__ a, int64_t ld, int64_t rows,
                                                               int cols, float* __restrict__ part) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * kColsumRows, r1 = min(rows, r0 + kColsumRows);
  if constexpr (WIDE) {
    __shared__ float red[8][32][9];
    const int c = blockIdx.x * 256 + tx * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c + 8 <= cols) {
      auto add = [&](const uint4& u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
      };
      int64_t r = r0 + ty;
      for (; r + 24 < r1; r += 32) {
        const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(a + r * ld + c));
        const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(a + (r + 8) * ld + c));
        const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(a + (r + 16) * ld + c));
        const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(a + (r + 24) * ld + c));
        add(u0); add(u1); add(u2); add(u3);
      }
      for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(a + r * ld + c)));
    } else {
      for (int j = 0; j < 8; ++j)
        if (c + j < cols)
          for (int64_t r = r0 + ty; r < r1; r += 8) acc[j] += __bfloat162float(a[r * ld + c + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty][tx][j] = acc[j];
    __syncthreads();
    // 256 threads: thread t sums column (t) of the block over the 8 row lanes
    const int cc = threadIdx.x;
    if (blockIdx.x * 256 + cc < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][cc >> 3][cc & 7];
      part[(int64_t)blockIdx.y * cols + blockIdx.x * 256 + cc] = t;
    }
  } else {
    __shared__ float2 red[8][33];
    const int c = blockIdx.x * 64 + tx * 2;
    float2 s = make_float2(0.f, 0.f);
    if (c + 1 < cols) {
      for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + r * ld + c));
        s.x += f.x; s.y += f.y;
      }
    } else if (c < cols) {
      for (int64_t r = r0 + ty; r < r1; r += 8) s.x += __bfloat162float(a[r * ld + c]);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
      float2 t = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { t.x += red[i][tx].x; t.y += red[i][tx].y; }
      float* p = part + (int64_t)blockIdx.y * cols + c;
      p[0] = t.x;
      if (c + 1 < cols) p[1] = t.y;
    }
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int nparts, int cols,
                                                           float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int k = 0; k < nparts; ++k) acc += part[(int64_t)k * cols + c];
  out[c] = acc;
}

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_dropout(int dtype, const void* x, void* y, int64_t n, float p, const void* rng_state, uint32_t site,
                           void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(x && y && rng_state && (((uintptr_t)x | (uintptr_t)y) % 16 == 0), "tcd_dropout: null or misaligned pointer");
  TCD_REQUIRE(p >= 0.f && p < 1.f, "tcd_dropout: p must be in [0, 1)");
  const uint32_t thr = drop_threshold(p);
  const float rk = 1.0f / (1.0f - p);
  const uint64_t* st = (const uint64_t*)rng_state;
  if (dtype == TCD_BF16)
    dropout_kernel<__nv_bfloat16><<<ceil_div(n, 2048), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n,
                                                                                   thr, rk, st, site);
  else if (dtype == TCD_F32)
    dropout_kernel<float><<<ceil_div(n, 1024), 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, n, thr, rk, st, site);
  else { set_error("tcd_dropout: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("dropout");
}

extern "C" int tcd_dropout_mask_attention(float* out, int samples, int heads, int Lq, int Lk, float p, const void* rng_state,
                                          uint32_t site, void* stream) {
  const int64_t rows = (int64_t)samples * heads * Lq;
  if (rows == 0 || Lk == 0) return TCD_OK;
  TCD_REQUIRE(out && rng_state && p >= 0.f && p < 1.f, "tcd_dropout_mask_attention: bad arguments");
  dropout_mask_attention_kernel<<<ceil_div(rows * Lk, 256), 256, 0, as_stream(stream)>>>(out, rows, Lk, drop_threshold(p),
                                                                                      1.0f / (1.0f - p),
                                                                                      (const uint64_t*)rng_state, site);
  return check_launch("dropout_mask_attention");
}

extern "C" int tcd_act_forward_bf16(int act, const void* z, void* y, int64_t n, float dropout_p, const void* rng_state,
                                    uint32_t site, void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(z && y && (((uintptr_t)z | (uintptr_t)y) % 16 == 0), "tcd_act_forward_bf16: null or misaligned pointer");
  TCD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || rng_state), "tcd_act_forward_bf16: bad dropout arguments");
  if (act == TCD_ACT_GELU)
    act16_kernel<false, true><<<ceil_div(n, 2048), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)z, nullptr, (__nv_bfloat16*)y, n, act,
                                                                               make_drop(dropout_p, rng_state, site));
  else
    act16_kernel<false, false><<<ceil_div(n, 2048), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)z, nullptr, (__nv_bfloat16*)y, n, act,
                                                                                make_drop(dropout_p, rng_state, site));
  return check_launch("act_forward_bf16");
}

extern "C" int tcd_act_backward_bf16(int act, const void* z, const void* dy, void* dx, int64_t n, float dropout_p,
                                     const void* rng_state, uint32_t site, void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(z && dy && dx && (((uintptr_t)z | (uintptr_t)dy | (uintptr_t)dx) % 16 == 0),
              "tcd_act_backward_bf16: null or misaligned pointer");
  if (act == TCD_ACT_GELU)
    act16_kernel<true, true><<<ceil_div(n, 2048), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)z, (const __nv_bfloat16*)dy,
                                                                              (__nv_bfloat16*)dx, n, act,
                                                                              make_drop(dropout_p, rng_state, site));
  else
    act16_kernel<true, false><<<ceil_div(n, 2048), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)z, (const __nv_bfloat16*)dy,
                                                                               (__nv_bfloat16*)dx, n, act,
                                                                               make_drop(dropout_p, rng_state, site));
  return check_launch("act_backward_bf16");
}

template <typename XT, typename DYT>
static int launch_lnb(const void* x, const float* gamma, const void* dy, const void* dyrot, const float* rc, const float* rs,
                      int tps, float eps, const void* dres, void* dx, float* gp, float* bp, int64_t rows, int D, int blocks,
                      DropArg dr, cudaStream_t st) {
#define TCD_LNB(NV)                                                                                                       \
  layernorm_backward_mixed_kernel<XT, DYT, NV><<<blocks, 256, 0, st>>>((const XT*)x, gamma, (const DYT*)dy, (const DYT*)dyrot, \
                                                                       rc, rs, tps, eps, (const XT*)dres, (XT*)dx, gp, bp, rows, dr)
  switch (D / 128) {
    case 1: TCD_LNB(1); break;
    case 2: TCD_LNB(2); break;
    case 4: TCD_LNB(4); break;
    case 8: TCD_LNB(8); break;
    default: set_error("tcd_layernorm_backward_mixed: D must be 128, 256, 512 or 1024"); return TCD_ERR_INVALID;
  }
#undef TCD_LNB
  return check_launch("layernorm_backward_mixed");
}

extern "C" int64_t tcd_layernorm_backward_mixed_partials(int64_t rows) { return tcd_layernorm_backward_partials(rows) / 8; }

extern "C" int tcd_layernorm_backward_mixed(int x_dtype, int dy_dtype, const void* x, const float* gamma, const void* dy,
                                            const void* dy_rot, const float* rot_cos, const float* rot_sin,
                                            int tokens_per_sample, float eps, const void* dres, void* dx, float* dgamma_part,
                                            float* dbeta_part, int64_t rows, int D, float x_dropout_p, const void* rng_state,
                                            uint32_t site, void* stream) {
  if (rows == 0) return TCD_OK;
  TCD_REQUIRE(x && gamma && (dy || dy_rot) && dx && dgamma_part && dbeta_part, "tcd_layernorm_backward_mixed: null pointer");
  TCD_REQUIRE(!dy_rot || (rot_cos && rot_sin && tokens_per_sample > 0), "tcd_layernorm_backward_mixed: rotary table missing");
  TCD_REQUIRE(x_dropout_p >= 0.f && x_dropout_p < 1.f && (x_dropout_p == 0.f || rng_state), "tcd_layernorm_backward_mixed: bad dropout arguments");
  const int blocks = (int)(tcd_layernorm_backward_partials(rows) / 8);
  cudaStream_t st = as_stream(stream);
  const DropArg dr = make_drop(x_dropout_p, rng_state, site);
#define TCD_LNB_CALL(XT, DYT)                                                                                             \
  return launch_lnb<XT, DYT>(x, gamma, dy, dy_rot, rot_cos, rot_sin, tokens_per_sample, eps, dres, dx, dgamma_part, dbeta_part, \
                             rows, D, blocks, dr, st)
  if (x_dtype == TCD_F32 && dy_dtype == TCD_BF16) TCD_LNB_CALL(float, __nv_bfloat16);
  if (x_dtype == TCD_BF16 && dy_dtype == TCD_BF16) TCD_LNB_CALL(__nv_bfloat16, __nv_bfloat16);
  if (x_dtype == TCD_F32 && dy_dtype == TCD_F32) TCD_LNB_CALL(float, float);
#undef TCD_LNB_CALL
  set_error("tcd_layernorm_backward_mixed: unsupported dtypes x=%d dy=%d", x_dtype, dy_dtype);
  return TCD_ERR_INVALID;
}

extern "C" int tcd_layernorm_bf16(const void* x, const float* gamma, const float* beta, float eps, void* y, int64_t rows, int D,
                                  float x_dropout_p, const void* rng_state, uint32_t site, void* stream) {
  if (rows == 0) return TCD_OK;
  TCD_REQUIRE(x && gamma && beta && y, "tcd_layernorm_bf16: null pointer");
  cudaStream_t st = as_stream(stream);
  const int blocks = ceil_div(rows, 8);
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x;
  __nv_bfloat16* yp = (__nv_bfloat16*)y;
  TCD_REQUIRE(x_dropout_p >= 0.f && x_dropout_p < 1.f && (x_dropout_p == 0.f || rng_state), "tcd_layernorm_bf16: bad dropout arguments");
  const DropArg dr = make_drop(x_dropout_p, rng_state, site);
  switch (D / 128) {
    case 1: layernorm16_kernel<1><<<blocks, 256, 0, st>>>(xp, gamma, beta, eps, yp, rows, dr); break;
    case 2: layernorm16_kernel<2><<<blocks, 256, 0, st>>>(xp, gamma, beta, eps, yp, rows, dr); break;
    case 4: layernorm16_kernel<4><<<blocks, 256, 0, st>>>(xp, gamma, beta, eps, yp, rows, dr); break;
    case 8: layernorm16_kernel<8><<<blocks, 256, 0, st>>>(xp, gamma, beta, eps, yp, rows, dr); break;
    default: set_error("tcd_layernorm_bf16: D must be 128, 256, 512 or 1024"); return TCD_ERR_INVALID;
  }
  return check_launch("layernorm_bf16");
}

extern "C" int64_t tcd_film_backward_workspace_floats(int samples, int L, int D) {
  return (int64_t)samples * ((L + kFilmChunk - 1) / kFilmChunk) * 2 * D;
}

extern "C" int tcd_film_backward_bf16(const float* dout, const void* v, const float* film, int64_t film_ld, int64_t film_off,
                                      void* dv, float* dfilm, int64_t dfilm_ld, int64_t dfilm_off, float* workspace,
                                      int samples, int L, int D, float v_dropout_p, const void* rng_state, uint32_t site,
                                      void* stream) {
  if (samples == 0 || L == 0) return TCD_OK;
  TCD_REQUIRE(dout && v && film && dv && dfilm && workspace, "tcd_film_backward_bf16: null pointer");
  TCD_REQUIRE(D % 4 == 0 && D <= 1024 && samples <= 65535 && film_off % 4 == 0 && film_ld % 4 == 0,
              "tcd_film_backward_bf16: D must be a multiple of 4 (<= 1024), film offsets 16-byte aligned");
  const int chunks = (L + kFilmChunk - 1) / kFilmChunk;
  cudaStream_t st = as_stream(stream);
  film_backward16_kernel<<<dim3(chunks, samples), D / 4, 0, st>>>(dout, (const __nv_bfloat16*)v, film, film_ld, film_off,
                                                                 (__nv_bfloat16*)dv, workspace, L, D, chunks,
                                                                 make_drop(v_dropout_p, rng_state, site));
  int rc = check_launch("film_backward_bf16");
  if (rc) return rc;
  film_reduce_kernel<<<dim3(ceil_div(2 * D, 256), samples), 256, 0, st>>>(workspace, dfilm, dfilm_ld, dfilm_off, chunks, 2 * D);
  return check_launch("film_reduce");
}

extern "C" int tcd_film_residual_bf16(const float* x, const void* v, const float* film, int64_t film_ld, int64_t film_off,
                                      float* out, int64_t rows, int L, int D, float v_dropout_p, const void* rng_state,
                                      uint32_t site, void* stream) {
  if (rows == 0) return TCD_OK;
  TCD_REQUIRE(x && v && out && D % 4 == 0 && L > 0, "tcd_film_residual_bf16: bad arguments");
  TCD_REQUIRE(!film || (film_off % 4 == 0 && film_ld % 4 == 0), "tcd_film_residual_bf16: film offsets must be 16-byte aligned");
  TCD_REQUIRE(v_dropout_p >= 0.f && v_dropout_p < 1.f && (v_dropout_p == 0.f || rng_state), "tcd_film_residual_bf16: bad dropout arguments");
  const int64_t n4 = rows * (D / 4);
  film_residual16_kernel<<<ceil_div(n4, 256), 256, 0, as_stream(stream)>>>(x, (const __nv_bfloat16*)v, film, film_ld, film_off, out,
                                                                          rows, L, D, make_drop(v_dropout_p, rng_state, site));
  return check_launch("film_residual_bf16");
}

extern "C" int64_t tcd_colsum_bf16_workspace_floats(int64_t rows, int cols) {
  return ((rows + kColsumRows - 1) / kColsumRows) * (int64_t)cols;
}

extern "C" int tcd_colsum_bf16(const void* a, int64_t ld, int64_t rows, int cols, float* out, float* workspace, void* stream) {
  TCD_REQUIRE(rows >= 0 && cols > 0, "tcd_colsum_bf16: bad shape");
  TCD_REQUIRE(a && out && workspace && ld % 2 == 0 && (uintptr_t)a % 4 == 0, "tcd_colsum_bf16: bad arguments");
  const int nparts = (int)((rows + kColsumRows - 1) / kColsumRows);
  TCD_REQUIRE(nparts <= 65535, "tcd_colsum_bf16: too many rows");
  cudaStream_t st = as_stream(stream);
  if (nparts > 0) {
    if (ld % 8 == 0 && (uintptr_t)a % 16 == 0)
      colsum16_partial_kernel<true><<<dim3(ceil_div(cols, 256), nparts), 256, 0, st>>>((const __nv_bfloat16*)a, ld, rows, cols, workspace);
    else
      colsum16_partial_kernel<false><<<dim3(ceil_div(cols, 64), nparts), 256, 0, st>>>((const __nv_bfloat16*)a, ld, rows, cols, workspace);
    int rc = check_launch("colsum_bf16");
    if (rc) return rc;
  }
  colsum_final_kernel<<<ceil_div(cols, 256), 256, 0, st>>>(workspace, nparts, cols, out);
  return check_launch("colsum_final");
}
