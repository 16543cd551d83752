// Backward-pass primitives of the training step (fp32 gradients; GEMM operands are cast to the operand dtype on the
// fly).  These kernels are the reverse-mode counterparts of the forward building blocks in norm.cu / gemm_*.cu:
//   tcd_cast_transpose     fp32 (R,C) -> operand dtype (C,R): operands of dgrad / wgrad GEMMs (tcd_gemm computes A W^T)
//   tcd_group_colsum       per-group column sums: bias gradients, LayerNorm dgamma/dbeta partials, FiLM dscale/dshift
//   tcd_act_forward/backward   ReLU / GELU(erf) / Mish / SiLU with the pre-activation kept for the backward
//   tcd_layernorm_backward dx and per-warp partial dgamma/dbeta
//   tcd_film_backward      out = x + (1+scale) v + shift : dv, dscale, dshift (dx = dout)
#include "common.cuh"

namespace tcd {

// ---------------------------------------------------------------------------------------- cast + transpose
template <typename T>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ src, int64_t src_ld,
                                                             T* __restrict__ dst, int64_t dst_ld, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? __ldg(src + r * src_ld + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;               // dst row = source column
    if (c < cols && r < rows) dst[c * dst_ld + r] = Conv<T>::to(tile[tx][ty + i]);
  }
}

// ---------------------------------------------------------------------------------------- all weight packs of a step
// One launch refreshes EVERY bf16 operand copy the training step needs (W for the forward / wgrad, W^T for dgrad) from the
// fp32 master parameters: segment table in device memory, blockIdx.y = segment, 32 x 32 tiles read once and written twice.
// (r01 profile of the c3 step: 129 convert_pad + 105 cast_transpose launches = 2.3 ms for 450 MB of traffic, i.e. launch
// bound; one launch moves them in ~0.1 ms.)  Pad rows / columns of the destinations are never written (allocated zeroed).
struct PackSeg {
  const float* src;        // (rows, cols) fp32, pitch src_ld
  __nv_bfloat16* w;        // (rows, w_ld) bf16 copy
  __nv_bfloat16* wt;       // (cols, wt_ld) bf16 transposed copy, or NULL
  int64_t src_ld, w_ld, wt_ld;
  int rows, cols;
};
static_assert(sizeof(PackSeg) == 56, "PackSeg layout is part of the C-ABI (tcd_pack_weights)");

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackSeg* __restrict__ segs) {
  __shared__ float tile[32][33];
  const PackSeg sg = segs[blockIdx.y];
  const int tc = (sg.cols + 31) / 32, tr = (sg.rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int t = blockIdx.x; t < tc * tr; t += gridDim.x) {
    const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const int r = r0 + ty + i, c = c0 + tx;
      const float v = (r < sg.rows && c < sg.cols) ? __ldg(sg.src + (int64_t)r * sg.src_ld + c) : 0.f;
      tile[ty + i][tx] = v;
      if (r < sg.rows && c < sg.cols) sg.w[(int64_t)r * sg.w_ld + c] = __float2bfloat16_rn(v);
    }
    if (sg.wt != nullptr) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;                 // destination row = source column
        if (c < sg.cols && r < sg.rows) sg.wt[(int64_t)c * sg.wt_ld + r] = __float2bfloat16_rn(tile[tx][ty + i]);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------- grouped column sums
// out[g, c] (+)= sum over the rows of group g of a[row, c] * (b ? b[row, c] : 1); one block per (32 columns, group)
__global__ void __launch_bounds__(256) group_colsum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int64_t ld, int64_t rows_per_group, int cols,
                                                           float* __restrict__ out, int64_t out_ld, int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t g = blockIdx.y;
  const float* pa = a + g * rows_per_group * ld;
  const float* pb = b ? b + g * rows_per_group * ld : nullptr;
  float s = 0.f;
  if (c < cols) {
    for (int64_t r = ty; r < rows_per_group; r += 8) {
      float v = __ldg(pa + r * ld + c);
      if (pb) v *= __ldg(pb + r * ld + c);
      s += v;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    float* o = out + g * out_ld + c;
    *o = accumulate ? *o + t : t;
  }
}

// ---------------------------------------------------------------------------------------- activations
__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case TCD_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case TCD_ACT_GELU: {   // d/dz [0.5 z (1 + erf(z/sqrt2))] = Phi(z) + z phi(z)
      const float phi = 0.3989422804014327f * expf(-0.5f * z * z);
      return 0.5f * (1.0f + erff(z * 0.70710678118654752440f)) + z * phi;
    }
    case TCD_ACT_MISH: {   // z tanh(sp(z)):  tanh(sp) + z (1 - tanh^2(sp)) sigmoid(z)
      const float sp = softplus_t(z), th = tanhf(sp);
      const float sg = 1.0f / (1.0f + expf(-z));
      return th + z * (1.0f - th * th) * sg;
    }
    case TCD_ACT_SILU: {
      const float sg = 1.0f / (1.0f + expf(-z));
      return sg * (1.0f + z * (1.0f - sg));
    }
    default: return 1.f;
  }
}
__global__ void __launch_bounds__(256) act_forward_kernel(const float* __restrict__ z, float* __restrict__ y, int64_t n, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = apply_act(z[i], act);
}
__global__ void __launch_bounds__(256) act_backward_kernel(const float* __restrict__ z, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int64_t n, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * act_grad(z[i], act);
}

// ---------------------------------------------------------------------------------------- LayerNorm backward
// warp per row (grid-stride over rows); lane owns columns {4*lane + 128*k}; per-warp partial dgamma/dbeta rows
template <int NV>
__global__ void __launch_bounds__(256) layernorm_backward_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ dy, float eps, float* __restrict__ dx,
                                                                 float* __restrict__ dgamma_part, float* __restrict__ dbeta_part,
                                                                 int64_t rows) {
  constexpr int D = 128 * NV;
  constexpr float invD = 1.0f / D;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    g[k] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp; row < rows; row += nwarps) {
    float4 xv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      xv[k] = reinterpret_cast<const float4*>(x + row * D)[lane + 32 * k];
      dv[k] = reinterpret_cast<const float4*>(dy + row * D)[lane + 32 * k];
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
    float s1 = 0.f, s2 = 0.f;   // sum(dxhat), sum(dxhat * xhat)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      xv[k].x *= rstd; xv[k].y *= rstd; xv[k].z *= rstd; xv[k].w *= rstd;        // xhat
      ag[k].x += dv[k].x * xv[k].x; ag[k].y += dv[k].y * xv[k].y; ag[k].z += dv[k].z * xv[k].z; ag[k].w += dv[k].w * xv[k].w;
      ab[k].x += dv[k].x; ab[k].y += dv[k].y; ab[k].z += dv[k].z; ab[k].w += dv[k].w;
      dv[k].x *= g[k].x; dv[k].y *= g[k].y; dv[k].z *= g[k].z; dv[k].w *= g[k].w;  // dxhat
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
      reinterpret_cast<float4*>(dx + row * D)[lane + 32 * k] = o;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    reinterpret_cast<float4*>(dgamma_part + warp * D)[lane + 32 * k] = ag[k];
    reinterpret_cast<float4*>(dbeta_part + warp * D)[lane + 32 * k] = ab[k];
  }
}

// ---------------------------------------------------------------------------------------- FiLM backward
// out = x + (1 + scale[b]) * v + shift[b]: dv = (1 + scale) dout; dscale[b] = sum_rows dout*v; dshift[b] = sum_rows dout.
// block = (sample, 128-column tile); thread = column, rows walked sequentially (coalesced across threads)
__global__ void __launch_bounds__(128) film_backward_kernel(const float* __restrict__ dout, const float* __restrict__ v,
                                                            const float* __restrict__ film, int64_t film_ld, int64_t film_off,
                                                            float* __restrict__ dv, float* __restrict__ dfilm, int64_t dfilm_ld,
                                                            int64_t dfilm_off, int L, int D) {
  const int b = blockIdx.y, c = blockIdx.x * 128 + threadIdx.x;
  if (c >= D) return;
  const float sc = 1.0f + __ldg(film + b * film_ld + film_off + c);
  float ds = 0.f, dsh = 0.f;
  const int64_t base = (int64_t)b * L * D + c;
  for (int r = 0; r < L; ++r) {
    const float g = dout[base + (int64_t)r * D];
    const float vv = v[base + (int64_t)r * D];
    dv[base + (int64_t)r * D] = sc * g;
    ds += g * vv;
    dsh += g;
  }
  dfilm[b * dfilm_ld + dfilm_off + c] = ds;
  dfilm[b * dfilm_ld + dfilm_off + D + c] = dsh;
}

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_cast_transpose(int dtype, const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                                  int64_t cols, void* stream) {
  TCD_REQUIRE(rows >= 0 && cols >= 0, "tcd_cast_transpose: bad shape");
  if (rows == 0 || cols == 0) return TCD_OK;
  TCD_REQUIRE(src && dst && src_ld >= cols && dst_ld >= rows, "tcd_cast_transpose: bad arguments");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  TCD_REQUIRE(grid.y <= 65535, "tcd_cast_transpose: too many rows (%lld)", (long long)rows);
  if (dtype == TCD_F32) cast_transpose_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(src, src_ld, (float*)dst, dst_ld, rows, cols);
  else if (dtype == TCD_BF16) cast_transpose_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(src, src_ld, (__nv_bfloat16*)dst, dst_ld, rows, cols);
  else { set_error("tcd_cast_transpose: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("cast_transpose");
}

extern "C" int tcd_pack_weights(const void* segments, int n_segments, int blocks_per_segment, void* stream) {
  TCD_REQUIRE(n_segments >= 0 && n_segments <= 65535 && blocks_per_segment > 0, "tcd_pack_weights: bad arguments");
  if (n_segments == 0) return TCD_OK;
  TCD_REQUIRE(segments != nullptr, "tcd_pack_weights: null segment table");
  dim3 grid((unsigned)blocks_per_segment, (unsigned)n_segments);
  pack_weights_kernel<<<grid, 256, 0, as_stream(stream)>>>((const PackSeg*)segments);
  return check_launch("pack_weights");
}

extern "C" int tcd_group_colsum(const float* a, const float* b, int64_t ld, int64_t groups, int64_t rows_per_group, int cols,
                                float* out, int64_t out_ld, int accumulate, void* stream) {
  TCD_REQUIRE(groups >= 0 && rows_per_group >= 0 && cols > 0, "tcd_group_colsum: bad shape");
  if (groups == 0) return TCD_OK;
  TCD_REQUIRE(a && out && groups <= 65535, "tcd_group_colsum: bad arguments (groups=%lld)", (long long)groups);
  dim3 grid(ceil_div(cols, 32), (unsigned)groups);
  group_colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, ld, rows_per_group, cols, out, out_ld, accumulate);
  return check_launch("group_colsum");
}

extern "C" int tcd_act_forward(int act, const float* z, float* y, int64_t n, void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(z && y, "tcd_act_forward: null pointer");
  act_forward_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(z, y, n, act);
  return check_launch("act_forward");
}

extern "C" int tcd_act_backward(int act, const float* z, const float* dy, float* dx, int64_t n, void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(z && dy && dx, "tcd_act_backward: null pointer");
  act_backward_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(z, dy, dx, n, act);
  return check_launch("act_backward");
}

extern "C" int64_t tcd_layernorm_backward_partials(int64_t rows) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = 148 * 4;
  return (blocks < cap ? blocks : cap) * 8;
}

extern "C" int tcd_layernorm_backward(const float* x, const float* gamma, const float* dy, float eps, float* dx,
                                      float* dgamma_part, float* dbeta_part, int64_t rows, int D, void* stream) {
  if (rows == 0) return TCD_OK;
  TCD_REQUIRE(x && gamma && dy && dx && dgamma_part && dbeta_part, "tcd_layernorm_backward: null pointer");
  const int blocks = (int)(tcd_layernorm_backward_partials(rows) / 8);
  cudaStream_t st = as_stream(stream);
  switch (D / 128) {
    case 1: layernorm_backward_kernel<1><<<blocks, 256, 0, st>>>(x, gamma, dy, eps, dx, dgamma_part, dbeta_part, rows); break;
    case 2: layernorm_backward_kernel<2><<<blocks, 256, 0, st>>>(x, gamma, dy, eps, dx, dgamma_part, dbeta_part, rows); break;
    case 4: layernorm_backward_kernel<4><<<blocks, 256, 0, st>>>(x, gamma, dy, eps, dx, dgamma_part, dbeta_part, rows); break;
    case 8: layernorm_backward_kernel<8><<<blocks, 256, 0, st>>>(x, gamma, dy, eps, dx, dgamma_part, dbeta_part, rows); break;
    default: set_error("tcd_layernorm_backward: D must be 128, 256, 512 or 1024"); return TCD_ERR_INVALID;
  }
  return check_launch("layernorm_backward");
}

extern "C" int tcd_film_backward(const float* dout, const float* v, const float* film, int64_t film_ld, int64_t film_off,
                                 float* dv, float* dfilm, int64_t dfilm_ld, int64_t dfilm_off, int samples, int L, int D,
                                 void* stream) {
  if (samples == 0 || L == 0) return TCD_OK;
  TCD_REQUIRE(dout && v && film && dv && dfilm, "tcd_film_backward: null pointer");
  TCD_REQUIRE(samples <= 65535, "tcd_film_backward: too many samples");
  film_backward_kernel<<<dim3(ceil_div(D, 128), samples), 128, 0, as_stream(stream)>>>(dout, v, film, film_ld, film_off, dv,
                                                                                        dfilm, dfilm_ld, dfilm_off, L, D);
  return check_launch("film_backward");
}
