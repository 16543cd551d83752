// Row-wise fp32 kernels of the denoiser: LayerNorm (+rotary) operand producers, the fused
// FiLM/residual/LayerNorm block tail, and the small conditioning-path kernels.
// One warp owns one row of D = 128*NV floats held in registers as NV float4 per lane (coalesced
// 512-byte warp accesses); statistics are fp32 two-pass in registers.
// Reference: model/model.py:171-173,219-220,326-344,375,387-388,585-616;
//            model/rotary_embedding_torch.py:39-59,107-130; model/utils.py:41-48.
#include "common.cuh"
#include "tuning.cuh"

namespace tcd {

constexpr int kWarpsPerBlock = 8;

template <int NV>
struct Row {
  float4 v[NV];
};

template <int NV>
__device__ __forceinline__ void row_load(Row<NV>& r, const float* p, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) r.v[k] = reinterpret_cast<const float4*>(p)[lane + 32 * k];
}
template <int NV>
__device__ __forceinline__ void row_load(Row<NV>& r, const __nv_bfloat16* p, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    uint2 u = reinterpret_cast<const uint2*>(p)[lane + 32 * k];
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    r.v[k] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
}
template <int NV>
__device__ __forceinline__ void row_store(const Row<NV>& r, float* p, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) reinterpret_cast<float4*>(p)[lane + 32 * k] = r.v[k];
}
template <int NV>
__device__ __forceinline__ void row_store(const Row<NV>& r, __nv_bfloat16* p, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r.v[k].x, r.v[k].y), b = __floats2bfloat162_rn(r.v[k].z, r.v[k].w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(p)[lane + 32 * k] = u;
  }
}

// y = (x - mean) / sqrt(var + eps) * gamma + beta, biased variance (torch.nn.LayerNorm)
template <int NV>
__device__ __forceinline__ void row_layernorm(Row<NV>& r, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int lane) {
  constexpr float invD = 1.0f / (128.0f * NV);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += (r.v[k].x + r.v[k].y) + (r.v[k].z + r.v[k].w);
  const float mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float a = r.v[k].x - mean, b = r.v[k].y - mean, c = r.v[k].z - mean, d = r.v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * invD + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
    float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * k);
    r.v[k].x = (r.v[k].x - mean) * rstd * g.x + b.x;
    r.v[k].y = (r.v[k].y - mean) * rstd * g.y + b.y;
    r.v[k].z = (r.v[k].z - mean) * rstd * g.z + b.z;
    r.v[k].w = (r.v[k].w - mean) * rstd * g.w + b.w;
  }
}

// interleaved-pair rotation by the angle table row of this token: pairs (2i, 2i+1) share freq i
template <int NV>
__device__ __forceinline__ void row_rotary(const Row<NV>& in, Row<NV>& out, const float* __restrict__ cos_row,
                                           const float* __restrict__ sin_row, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float2 c = __ldg(reinterpret_cast<const float2*>(cos_row) + lane + 32 * k);
    float2 s = __ldg(reinterpret_cast<const float2*>(sin_row) + lane + 32 * k);
    out.v[k].x = in.v[k].x * c.x - in.v[k].y * s.x;
    out.v[k].y = in.v[k].y * c.x + in.v[k].x * s.x;
    out.v[k].z = in.v[k].z * c.y - in.v[k].w * s.y;
    out.v[k].w = in.v[k].w * c.y + in.v[k].z * s.y;
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) layernorm_rotary_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    T* __restrict__ out_plain, T* __restrict__ out_rot, const float* __restrict__ rot_cos,
    const float* __restrict__ rot_sin, int64_t rows, int tokens_per_sample) {
  constexpr int D = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  Row<NV> r;
  row_load<NV>(r, x + row * D, lane);
  row_layernorm<NV>(r, gamma, beta, eps, lane);
  if (out_plain) row_store<NV>(r, out_plain + row * D, lane);
  if (out_rot) {
    const int pos = (int)(row % tokens_per_sample);
    Row<NV> q;
    row_rotary<NV>(r, q, rot_cos + (int64_t)pos * (D / 2), rot_sin + (int64_t)pos * (D / 2), lane);
    row_store<NV>(q, out_rot + row * D, lane);
  }
}

// memory rows: tokens (n,S,D) followed by two time tokens per sample -> norm_cond -> plain + rotary
template <typename T, int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) build_memory_kernel(
    const float* __restrict__ tokens, const float* __restrict__ t_tokens, const float* __restrict__ gamma,
    const float* __restrict__ beta, T* __restrict__ mem_plain, T* __restrict__ mem_rot,
    const float* __restrict__ rot_cos, const float* __restrict__ rot_sin, int64_t rows, int S) {
  constexpr int D = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / (S + 2);
  const int s = (int)(row - b * (S + 2));
  const float* src = s < S ? tokens + (b * S + s) * D : t_tokens + (b * 2 + (s - S)) * D;
  Row<NV> r, q;
  row_load<NV>(r, src, lane);
  row_layernorm<NV>(r, gamma, beta, 1e-5f, lane);
  if (mem_plain) row_store<NV>(r, mem_plain + row * D, lane);
  if (mem_rot) {
    row_rotary<NV>(r, q, rot_cos + (int64_t)s * (D / 2), rot_sin + (int64_t)s * (D / 2), lane);
    row_store<NV>(q, mem_rot + row * D, lane);
  }
}

// one block per sample: where(keep, tokens, null) in place, mean over S, LayerNorm -> operand row
template <typename T>
__global__ void cond_pool_kernel(float* __restrict__ tokens, const float* __restrict__ null_embed,
                                 const uint8_t* __restrict__ keep, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, T* __restrict__ pooled, int S, int D) {
  __shared__ float s_red[32];
  __shared__ float s_stat[2];
  const int b = blockIdx.x;
  const bool kp = keep[b] != 0;
  float* tok = tokens + (int64_t)b * S * D;
  float vals[4];
  float mine = 0.f;
  int nmine = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    if (kp) {
      for (int s = 0; s < S; ++s) acc += tok[(int64_t)s * D + d];
    } else {
      for (int s = 0; s < S; ++s) {
        float v = __ldg(null_embed + (int64_t)s * D + d);
        tok[(int64_t)s * D + d] = v;
        acc += v;
      }
    }
    vals[nmine] = acc / (float)S;
    mine += vals[nmine];
    ++nmine;
  }
  // block mean
  float w = warp_sum(mine);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += s_red[i];
    s_stat[0] = t / (float)D;
  }
  __syncthreads();
  const float mean = s_stat[0];
  float q = 0.f;
  for (int i = 0; i < nmine; ++i) q += (vals[i] - mean) * (vals[i] - mean);
  w = warp_sum(q);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += s_red[i];
    s_stat[1] = 1.0f / sqrtf(t / (float)D + 1e-5f);
  }
  __syncthreads();
  const float rstd = s_stat[1];
  int i = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x, ++i)
    pooled[(int64_t)b * D + d] = Conv<T>::to((vals[i] - mean) * rstd * __ldg(gamma + d) + __ldg(beta + d));
}

template <typename T>
__global__ void time_cond_kernel(const float* __restrict__ t_lin, const float* __restrict__ cond_hidden,
                                 const float* __restrict__ null_hidden, const uint8_t* __restrict__ keep,
                                 float* __restrict__ t_out, T* __restrict__ mish_out, int n, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * D) return;
  int b = (int)(i / D), d = (int)(i - (int64_t)b * D);
  float t = t_lin[i] + (keep[b] ? cond_hidden[i] : __ldg(null_hidden + d));
  if (t_out) t_out[i] = t;
  if (mish_out) mish_out[i] = Conv<T>::to(act_mish(t));
}

// sampler: mish_out[s, j, :] = Mish(t_lin[s] + (j < B ? cond_hidden[j] : null_hidden_proj)) for j < 2B
template <typename T>
__global__ void sampler_time_cond_kernel(const float* __restrict__ t_lin, const float* __restrict__ ch_cond,
                                         const float* __restrict__ ch_uncond, T* __restrict__ mish_out, int steps,
                                         int B, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)2 * B * D;
  if (i >= per * steps) return;
  int s = (int)(i / per);
  int64_t r = i - (int64_t)s * per;
  int j = (int)(r / D), d = (int)(r - (int64_t)j * D);
  float t = t_lin[(int64_t)s * D + d] + (j < B ? ch_cond[(int64_t)j * D + d] : ch_uncond[d]);
  mish_out[i] = Conv<T>::to(act_mish(t));
}

template <typename T>
__global__ void time_embed_kernel(const int64_t* __restrict__ times, const float* __restrict__ table,
                                  T* __restrict__ out, int n, int D, int n_timestep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * D) return;
  int b = (int)(i / D), d = (int)(i - (int64_t)b * D);
  int64_t t = times[b];
  t = t < 0 ? 0 : (t >= n_timestep ? n_timestep - 1 : t);
  out[i] = Conv<T>::to(__ldg(table + t * D + d));
}

// standalone rotary (RotaryEmbedding.rotate_queries_or_keys): one thread per (row, pair)
__global__ void rotary_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ rot_cos,
                              const float* __restrict__ rot_sin, int64_t rows, int D, int tokens_per_sample) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = D / 2;
  if (i >= rows * half) return;
  int64_t row = i / half;
  int p = (int)(i - row * half);
  int pos = (int)(row % tokens_per_sample);
  float c = __ldg(rot_cos + (int64_t)pos * half + p), s = __ldg(rot_sin + (int64_t)pos * half + p);
  float2 v = reinterpret_cast<const float2*>(x)[i];
  reinterpret_cast<float2*>(out)[i] = make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
}

template <typename T>
__global__ void scatter_rows_kernel(const T* __restrict__ src, int64_t src_ld, int64_t src_batch_stride,
                                    T* __restrict__ dst, int64_t dst_ld, int64_t dst_batch_stride, int64_t dst_row0,
                                    int rows, int cols, int samples) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per = (int64_t)rows * cols;
  if (i >= per * samples) return;
  int b = (int)(i / per);
  int64_t r = i - (int64_t)b * per;
  int row = (int)(r / cols), c = (int)(r - (int64_t)row * cols);
  dst[b * dst_batch_stride + (dst_row0 + row) * dst_ld + c] = src[b * src_batch_stride + (int64_t)row * src_ld + c];
}

// x[b, s, d, c] = w[s, c] * value[b, s, d, c] + (1 - w[s, c]) * x[b, s, d, c]   (weights shared by batch and dancer)
__global__ void masked_blend_kernel(float* __restrict__ x, const float* __restrict__ value, const float* __restrict__ w,
                                    int64_t n, int S, int dn, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = (int)(i % C);
  int s = (int)((i / ((int64_t)C * dn)) % S);
  float wt = __ldg(w + (int64_t)s * C + c);
  if (wt == 0.f) return;
  x[i] = __fadd_rn(__fmul_rn(wt, __ldg(value + i)), __fmul_rn(__fsub_rn(1.0f, wt), x[i]));
}

template <typename T>
__global__ void convert_pad_kernel(const float* __restrict__ src, int64_t src_ld, T* __restrict__ dst,
                                   int64_t dst_ld, int64_t rows, int cols) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dst_ld) return;
  int64_t r = i / dst_ld;
  int c = (int)(i - r * dst_ld);
  dst[i] = Conv<T>::to(c < cols ? __ldg(src + r * src_ld + c) : 0.0f);
}

// fp32 -> bf16 when every pitch allows 16-byte accesses and cols is a multiple of 8: thread = 8 columns (two 16-byte loads,
// one 16-byte store), two independent vectors per iteration of a grid-stride loop.  (r02 launch list of the training step:
// the scalar kernel above took 152 us for a 96 000 x 512 gradient = 0.30 of HBM, 1.2 ms per step in 8 launches.)
__global__ void __launch_bounds__(256) convert_bf16_vec8_kernel(const float* __restrict__ src, int64_t src_ld, __nv_bfloat16* __restrict__ dst,
                                                                int64_t dst_ld, int64_t rows, int cols, int vec_per_row) {
  const int64_t nvec = rows * vec_per_row;                 // vec_per_row = dst_ld / 8 (columns past `cols` are written as zeros)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto load = [&](int64_t i, float4& a, float4& b) {
    const int64_t r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row) * 8;
    if (c < cols) {
      const float4* p = reinterpret_cast<const float4*>(src + r * src_ld + c);
      a = __ldg(p); b = __ldg(p + 1);
    } else {
      a = b = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return r * dst_ld + c;
  };
  auto store = [&](int64_t o, const float4& a, const float4& b) {
    __nv_bfloat162 h[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w), __floats2bfloat162_rn(b.x, b.y),
                           __floats2bfloat162_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(h);
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < nvec; i += 2 * stride) {
    float4 a0, b0, a1, b1;
    const int64_t o0 = load(i, a0, b0), o1 = load(i + stride, a1, b1);
    store(o0, a0, b0); store(o1, a1, b1);
  }
  if (i < nvec) {
    float4 a0, b0;
    const int64_t o0 = load(i, a0, b0);
    store(o0, a0, b0);
  }
}

// ---- FiLM + residual + LayerNorm block tail (model/model.py:171-173,327,334,339): persistent, software-pipelined ----------
// A one-row-per-warp kernel has its 3 KB of row loads in flight only at the start of a warp's short life
// (ncu r01: 46 % active warps, DRAM 55 % busy => latency-bound).  Here a resident grid strides over the rows and
// every warp issues the loads of its NEXT row (raw registers, converted only when used) before it touches the
// current one, so ~3 KB per warp stay in flight for the whole kernel.  Same arithmetic, same order => bit-identical.
template <typename TY, int NV>
struct RawRow;
template <int NV>
struct RawRow<float, NV> {
  float4 v[NV];
  __device__ __forceinline__ void load(const float* p, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = reinterpret_cast<const float4*>(p)[lane + 32 * k];
  }
  __device__ __forceinline__ void get(Row<NV>& r) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) r.v[k] = v[k];
  }
};
template <int NV>
struct RawRow<__nv_bfloat16, NV> {
  uint2 v[NV];
  __device__ __forceinline__ void load(const __nv_bfloat16* p, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = reinterpret_cast<const uint2*>(p)[lane + 32 * k];
  }
  __device__ __forceinline__ void get(Row<NV>& r) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      uint2 u = v[k];
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
      r.v[k] = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
    }
  }
};

// Bounded for 2 resident blocks per SM: ptxas then hoists the loop-invariant LayerNorm parameter vectors into
// registers (126 at D = 512); bounding for 3 blocks (80 registers, spills) measured 14 % slower.
template <typename T, typename TY, int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, NV <= 4 ? 2 : 1) film_residual_norm_pf_kernel(
    const float* x_in, float* x_out, const TY* __restrict__ y, const float* __restrict__ gin, const float* __restrict__ bin,
    float eps_in, const float* __restrict__ film, int64_t film_ld, int64_t film_off,
    const float* __restrict__ gnext, const float* __restrict__ bnext, float eps_next, T* __restrict__ out_plain,
    T* __restrict__ out_rot, const float* __restrict__ rot_cos, const float* __restrict__ rot_sin, int64_t rows,
    int tokens_per_sample) {
  constexpr int D = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock;
  int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  RawRow<TY, NV> ry;
  RawRow<float, NV> rx;
  ry.load(y + row * D, lane);
  rx.load(x_in + row * D, lane);
  while (true) {
    Row<NV> v, xr;
    ry.get(v);
    rx.get(xr);
    const int64_t nrow = row + stride;
    if (nrow < rows) {                                    // next row's loads fly while this row is processed
      ry.load(y + nrow * D, lane);
      rx.load(x_in + nrow * D, lane);
    }
    if (gin) row_layernorm<NV>(v, gin, bin, eps_in, lane);
    if (film) {
      const float* f = film + (row / tokens_per_sample) * film_ld + film_off;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float4 sc = __ldg(reinterpret_cast<const float4*>(f) + lane + 32 * k);
        float4 sh = __ldg(reinterpret_cast<const float4*>(f + D) + lane + 32 * k);
        xr.v[k].x += (sc.x + 1.0f) * v.v[k].x + sh.x;
        xr.v[k].y += (sc.y + 1.0f) * v.v[k].y + sh.y;
        xr.v[k].z += (sc.z + 1.0f) * v.v[k].z + sh.z;
        xr.v[k].w += (sc.w + 1.0f) * v.v[k].w + sh.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        xr.v[k].x += v.v[k].x; xr.v[k].y += v.v[k].y; xr.v[k].z += v.v[k].z; xr.v[k].w += v.v[k].w;
      }
    }
    if (x_out) row_store<NV>(xr, x_out + row * D, lane);   // NULL: the updated x is dead (only its LayerNorm is consumed)
    if (gnext) {
      row_layernorm<NV>(xr, gnext, bnext, eps_next, lane);
      if (out_plain) row_store<NV>(xr, out_plain + row * D, lane);
      if (out_rot) {
        const int pos = (int)(row % tokens_per_sample);
        Row<NV> q;
        row_rotary<NV>(xr, q, rot_cos + (int64_t)pos * (D / 2), rot_sin + (int64_t)pos * (D / 2), lane);
        row_store<NV>(q, out_rot + row * D, lane);
      }
    }
    if (nrow >= rows) break;
    row = nrow;
  }
}

// ---- round-2 variant: bulk-copy input ring + contiguous row chunks + register-resident parameters ---------------------------
// ncu r02 of the kernel above at the sampler's shape (96 000 rows of 512): 544 L1 sectors per row, of which only 96 are the
// row's own 3 KB — every warp re-reads the four LayerNorm vectors (8 KB), the FiLM scale / shift of its sample (4 KB) and the
// rotary row (2 KB) for each row it touches; L1 data pipe 60 % busy, DRAM 55 %, issue 44 %: nothing saturated, 0.67-0.75 of
// the HBM peak.  Here every warp owns a CONTIGUOUS chunk of rows, so
//   * the LayerNorm vectors are loaded once per warp and the FiLM vectors once per sample into registers (96 registers at
//     D = 512; one block of 12 warps per SM, 170 registers per thread): 160 sectors per row are left;
//   * the input rows arrive by cp.async.bulk in a per-warp ring of NS = 4 slots of [y row | x row] (an mbarrier per slot):
//     144 KB per SM in flight without spending registers on prefetch (the ring alone, with the parameter re-reads kept,
//     measured 3 % slower than the kernel above: profiles/r02_ab.md).
// Same arithmetic in the same order => bit-identical outputs.
__device__ __forceinline__ uint32_t nsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ring_wait(uint32_t bar, uint32_t parity) {
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
template <int NV>
struct Vec {
  float4 v[NV];
  __device__ __forceinline__ void load(const float* p, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(p) + lane + 32 * k);
  }
};
// row_layernorm with gamma / beta in registers (same operations in the same order)
template <int NV>
__device__ __forceinline__ void row_layernorm_r(Row<NV>& r, const Vec<NV>& gamma, const Vec<NV>& beta, float eps) {
  constexpr float invD = 1.0f / (128.0f * NV);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += (r.v[k].x + r.v[k].y) + (r.v[k].z + r.v[k].w);
  const float mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float a = r.v[k].x - mean, b = r.v[k].y - mean, c = r.v[k].z - mean, d = r.v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * invD + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    r.v[k].x = (r.v[k].x - mean) * rstd * gamma.v[k].x + beta.v[k].x;
    r.v[k].y = (r.v[k].y - mean) * rstd * gamma.v[k].y + beta.v[k].y;
    r.v[k].z = (r.v[k].z - mean) * rstd * gamma.v[k].z + beta.v[k].z;
    r.v[k].w = (r.v[k].w - mean) * rstd * gamma.v[k].w + beta.v[k].w;
  }
}

constexpr int kRcWarps = 12;

template <typename T, typename TY, int NV, int NS>
__global__ void __launch_bounds__(kRcWarps * 32, 1) film_residual_norm_rc_kernel(
    const float* x_in, float* x_out, const TY* __restrict__ y, const float* __restrict__ gin, const float* __restrict__ bin,
    float eps_in, const float* __restrict__ film, int64_t film_ld, int64_t film_off,
    const float* __restrict__ gnext, const float* __restrict__ bnext, float eps_next, T* __restrict__ out_plain,
    T* __restrict__ out_rot, const float* __restrict__ rot_cos, const float* __restrict__ rot_sin, int64_t rows,
    int tokens_per_sample) {
  constexpr int D = 128 * NV;
  constexpr uint32_t YB = D * sizeof(TY), XB = D * 4, SLOT = YB + XB;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t slots = nsmem_u32(ring_smem) + (uint32_t)warp * NS * SLOT;
  const uint32_t bars = nsmem_u32(ring_smem) + (uint32_t)kRcWarps * NS * SLOT + (uint32_t)warp * NS * 8;
  const uint8_t* slots_gen = ring_smem + (size_t)warp * NS * SLOT;
  const int64_t nwarps = (int64_t)gridDim.x * kRcWarps;
  const int64_t chunk = (rows + nwarps - 1) / nwarps;
  const int64_t r0 = ((int64_t)blockIdx.x * kRcWarps + warp) * chunk;
  const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
  if (r0 >= r1) return;                                    // whole warp; barriers are per warp, no block-wide sync below
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int64_t r = r0 + s;
      if (r < r1) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * s), "r"(SLOT) : "memory");
        bulk_g2s(slots + s * SLOT, y + r * D, YB, bars + 8u * s);
        bulk_g2s(slots + s * SLOT + YB, x_in + r * D, XB, bars + 8u * s);
      }
    }
  }
  Vec<NV> gi, bi, gn, bn, sc, sh;
  if (gin) { gi.load(gin, lane); bi.load(bin, lane); }
  if (gnext) { gn.load(gnext, lane); bn.load(bnext, lane); }
  int64_t sample = r0 / tokens_per_sample;
  int pos = (int)(r0 - sample * tokens_per_sample);
  if (film) {
    const float* f = film + sample * film_ld + film_off;
    sc.load(f, lane);
    sh.load(f + D, lane);
  }
  __syncwarp();
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t row = r0; row < r1; ++row) {
    ring_wait(bars + 8u * slot, phase);
    Row<NV> v, xr;
    row_load<NV>(v, reinterpret_cast<const TY*>(slots_gen + (size_t)slot * SLOT), lane);
    row_load<NV>(xr, reinterpret_cast<const float*>(slots_gen + (size_t)slot * SLOT + YB), lane);
    __syncwarp();                                          // every lane holds its part of the row: the slot is free
    const int64_t nrow = row + NS;
    if (lane == 0 && nrow < r1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * slot), "r"(SLOT) : "memory");
      bulk_g2s(slots + slot * SLOT, y + nrow * D, YB, bars + 8u * slot);
      bulk_g2s(slots + slot * SLOT + YB, x_in + nrow * D, XB, bars + 8u * slot);
    }
    if (gin) row_layernorm_r<NV>(v, gi, bi, eps_in);
    if (film) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        xr.v[k].x += (sc.v[k].x + 1.0f) * v.v[k].x + sh.v[k].x;
        xr.v[k].y += (sc.v[k].y + 1.0f) * v.v[k].y + sh.v[k].y;
        xr.v[k].z += (sc.v[k].z + 1.0f) * v.v[k].z + sh.v[k].z;
        xr.v[k].w += (sc.v[k].w + 1.0f) * v.v[k].w + sh.v[k].w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        xr.v[k].x += v.v[k].x; xr.v[k].y += v.v[k].y; xr.v[k].z += v.v[k].z; xr.v[k].w += v.v[k].w;
      }
    }
    if (x_out) row_store<NV>(xr, x_out + row * D, lane);
    if (gnext) {
      row_layernorm_r<NV>(xr, gn, bn, eps_next);
      if (out_plain) row_store<NV>(xr, out_plain + row * D, lane);
      if (out_rot) {
        Row<NV> q;
        row_rotary<NV>(xr, q, rot_cos + (int64_t)pos * (D / 2), rot_sin + (int64_t)pos * (D / 2), lane);
        row_store<NV>(q, out_rot + row * D, lane);
      }
    }
    if (++pos == tokens_per_sample) {                       // next sample: new FiLM vectors (warp-uniform)
      pos = 0;
      ++sample;
      if (film && row + 1 < r1) {
        const float* f = film + sample * film_ld + film_off;
        sc.load(f, lane);
        sh.load(f + D, lane);
      }
    }
    if (++slot == NS) { slot = 0; phase ^= 1u; }
  }
}

template <typename T, int NV>
static int launch_ln(const float* x, const float* g, const float* b, float eps, void* op, void* orot,
                     const float* rc, const float* rs, int64_t rows, int tps, cudaStream_t st) {
  layernorm_rotary_kernel<T, NV><<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, st>>>(
      x, g, b, eps, (T*)op, (T*)orot, rc, rs, rows, tps);
  return check_launch("layernorm_rotary");
}

int num_sms();

template <typename T, typename TY, int NV>
static int launch_frn_pf(const float* x_in, float* x_out, const void* y, const float* gi, const float* bi, float ei,
                         const float* film, int64_t fld, int64_t foff, const float* gn, const float* bn, float en, void* op,
                         void* orot, const float* rc, const float* rs, int64_t rows, int tps, cudaStream_t st) {
  static int resident = 0;                               // blocks that fit on the device at once (per instantiation)
  if (!resident) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, film_residual_norm_pf_kernel<T, TY, NV>,
                                                                  kWarpsPerBlock * 32, 0);
    if (e != cudaSuccess || per_sm < 1) { set_error("film_residual_norm: occupancy query: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    resident = per_sm * num_sms();
  }
  const int64_t want = ceil_div(rows, kWarpsPerBlock);
  const int grid = (int)(want < resident ? want : resident);
  film_residual_norm_pf_kernel<T, TY, NV><<<grid, kWarpsPerBlock * 32, 0, st>>>(
      x_in, x_out, (const TY*)y, gi, bi, ei, film, fld, foff, gn, bn, en, (T*)op, (T*)orot, rc, rs, rows, tps);
  return check_launch("film_residual_norm");
}

template <typename T, typename TY, int NV>
static int launch_frn_rc(const float* x_in, float* x_out, const void* y, const float* gi, const float* bi, float ei,
                         const float* film, int64_t fld, int64_t foff, const float* gn, const float* bn, float en, void* op,
                         void* orot, const float* rc, const float* rs, int64_t rows, int tps, cudaStream_t st) {
  constexpr int NS = 4;
  constexpr size_t smem = (size_t)kRcWarps * NS * (128 * NV * (sizeof(TY) + 4)) + kRcWarps * NS * 8;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(film_residual_norm_rc_kernel<T, TY, NV, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("film_residual_norm(rc): smem attribute: %s", cudaGetErrorString(e)); return TCD_ERR_CUDA; }
    configured = true;
  }
  const int64_t want = ceil_div(rows, kRcWarps * 4);       // at least 4 rows per warp
  const int grid = (int)(want < num_sms() ? want : num_sms());
  film_residual_norm_rc_kernel<T, TY, NV, NS><<<grid, kRcWarps * 32, smem, st>>>(
      x_in, x_out, (const TY*)y, gi, bi, ei, film, fld, foff, gn, bn, en, (T*)op, (T*)orot, rc, rs, rows, tps);
  return check_launch("film_residual_norm");
}

template <typename T, typename TY, int NV>
static int launch_frn(const float* x_in, float* x_out, const void* y, const float* gi, const float* bi, float ei, const float* film,
                      int64_t fld, int64_t foff, const float* gn, const float* bn, float en, void* op, void* orot,
                      const float* rc, const float* rs, int64_t rows, int tps, cudaStream_t st) {
#if TCD_TUNE_FRN_RC
  // big bf16-y launches at D <= 512 (the sampler's tails); the bulk copies need 16-byte aligned rows
  if constexpr (NV <= 4 && sizeof(TY) == 2) {
    if (rows >= 4096 && ((uintptr_t)x_in | (uintptr_t)y) % 16 == 0)
      return launch_frn_rc<T, TY, NV>(x_in, x_out, y, gi, bi, ei, film, fld, foff, gn, bn, en, op, orot, rc, rs, rows, tps, st);
  }
#endif
  return launch_frn_pf<T, TY, NV>(x_in, x_out, y, gi, bi, ei, film, fld, foff, gn, bn, en, op, orot, rc, rs, rows, tps, st);
}

template <typename T, int NV>
static int launch_mem(const float* tok, const float* tt, const float* g, const float* b, void* mp, void* mr,
                      const float* rc, const float* rs, int64_t rows, int S, cudaStream_t st) {
  build_memory_kernel<T, NV><<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, st>>>(
      tok, tt, g, b, (T*)mp, (T*)mr, rc, rs, rows, S);
  return check_launch("build_memory");
}

#define TCD_NV_SWITCH(NVVAL, CALL)                                                    \
  switch (NVVAL) {                                                                    \
    case 1: { constexpr int NV = 1; return CALL; }                                    \
    case 2: { constexpr int NV = 2; return CALL; }                                    \
    case 4: { constexpr int NV = 4; return CALL; }                                    \
    case 8: { constexpr int NV = 8; return CALL; }                                    \
    default: set_error("feature width D must be 128, 256, 512 or 1024"); return TCD_ERR_INVALID; \
  }

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_layernorm_rotary(int dtype, const float* x, const float* gamma, const float* beta, float eps,
                                    void* out_plain, void* out_rot, const float* rot_cos, const float* rot_sin,
                                    int64_t rows, int D, int tokens_per_sample, void* stream) {
  TCD_REQUIRE(x && gamma && beta && (out_plain || out_rot), "tcd_layernorm_rotary: null pointer");
  TCD_REQUIRE(!out_rot || (rot_cos && rot_sin && tokens_per_sample > 0), "tcd_layernorm_rotary: rotary table missing");
  TCD_REQUIRE(D % 128 == 0, "tcd_layernorm_rotary: D %% 128 != 0");
  if (rows == 0) return TCD_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == TCD_F32) {
    TCD_NV_SWITCH(D / 128, (launch_ln<float, NV>(x, gamma, beta, eps, out_plain, out_rot, rot_cos, rot_sin, rows, tokens_per_sample, st)))
  } else if (dtype == TCD_BF16) {
    TCD_NV_SWITCH(D / 128, (launch_ln<__nv_bfloat16, NV>(x, gamma, beta, eps, out_plain, out_rot, rot_cos, rot_sin, rows, tokens_per_sample, st)))
  }
  set_error("tcd_layernorm_rotary: bad dtype %d", dtype);
  return TCD_ERR_INVALID;
}

extern "C" int tcd_film_residual_norm(int dtype, const float* x_in, float* x_out, const void* y, int y_dtype, const float* ln_in_gamma,
                                      const float* ln_in_beta, float ln_in_eps, const float* film, int64_t film_ld,
                                      int64_t film_off, const float* next_gamma, const float* next_beta,
                                      float next_eps, void* out_plain, void* out_rot, const float* rot_cos,
                                      const float* rot_sin, int64_t rows, int D, int tokens_per_sample,
                                      void* stream) {
  TCD_REQUIRE(x_in && y, "tcd_film_residual_norm: null pointer");
  TCD_REQUIRE(x_out || next_gamma, "tcd_film_residual_norm: x_out == NULL needs a next LayerNorm output");
  TCD_REQUIRE((ln_in_gamma == nullptr) == (ln_in_beta == nullptr), "tcd_film_residual_norm: inner LN params");
  TCD_REQUIRE(!next_gamma || next_beta, "tcd_film_residual_norm: next LN params");
  TCD_REQUIRE(!out_rot || (rot_cos && rot_sin), "tcd_film_residual_norm: rotary table missing");
  TCD_REQUIRE(tokens_per_sample > 0 && D % 128 == 0, "tcd_film_residual_norm: bad shape");
  TCD_REQUIRE(!film || (film_ld % 4 == 0 && film_off % 4 == 0), "tcd_film_residual_norm: film alignment");
  if (rows == 0) return TCD_OK;
  cudaStream_t st = as_stream(stream);
#define FRN_ARGS x_in, x_out, y, ln_in_gamma, ln_in_beta, ln_in_eps, film, film_ld, film_off, next_gamma, next_beta, next_eps, \
                 out_plain, out_rot, rot_cos, rot_sin, rows, tokens_per_sample, st
  if (dtype == TCD_F32 && y_dtype == TCD_F32) {
    TCD_NV_SWITCH(D / 128, (launch_frn<float, float, NV>(FRN_ARGS)))
  } else if (dtype == TCD_BF16 && y_dtype == TCD_F32) {
    TCD_NV_SWITCH(D / 128, (launch_frn<__nv_bfloat16, float, NV>(FRN_ARGS)))
  } else if (dtype == TCD_BF16 && y_dtype == TCD_BF16) {
    TCD_NV_SWITCH(D / 128, (launch_frn<__nv_bfloat16, __nv_bfloat16, NV>(FRN_ARGS)))
  }
#undef FRN_ARGS
  set_error("tcd_film_residual_norm: unsupported dtype pair (%d, %d)", dtype, y_dtype);
  return TCD_ERR_INVALID;
}

extern "C" int tcd_build_memory(int dtype, const float* tokens, const float* t_tokens, const float* gamma,
                                const float* beta, void* mem_plain, void* mem_rot, const float* rot_cos,
                                const float* rot_sin, int n, int S, int D, void* stream) {
  TCD_REQUIRE(tokens && t_tokens && gamma && beta && (mem_plain || mem_rot), "tcd_build_memory: null pointer");
  TCD_REQUIRE(!mem_rot || (rot_cos && rot_sin), "tcd_build_memory: rotary table missing");
  if (n == 0) return TCD_OK;
  cudaStream_t st = as_stream(stream);
  const int64_t rows = (int64_t)n * (S + 2);
  if (dtype == TCD_F32) {
    TCD_NV_SWITCH(D / 128, (launch_mem<float, NV>(tokens, t_tokens, gamma, beta, mem_plain, mem_rot, rot_cos, rot_sin, rows, S, st)))
  } else if (dtype == TCD_BF16) {
    TCD_NV_SWITCH(D / 128, (launch_mem<__nv_bfloat16, NV>(tokens, t_tokens, gamma, beta, mem_plain, mem_rot, rot_cos, rot_sin, rows, S, st)))
  }
  set_error("tcd_build_memory: bad dtype %d", dtype);
  return TCD_ERR_INVALID;
}

extern "C" int tcd_cond_pool(int dtype, float* tokens, const float* null_embed, const uint8_t* keep,
                             const float* gamma, const float* beta, void* pooled_ln, int n, int S, int D,
                             void* stream) {
  TCD_REQUIRE(tokens && null_embed && keep && gamma && beta && pooled_ln, "tcd_cond_pool: null pointer");
  TCD_REQUIRE(D <= 1024 && D % 32 == 0, "tcd_cond_pool: D must be <= 1024 and a multiple of 32");
  if (n == 0) return TCD_OK;
  const int threads = D >= 256 ? 256 : D;  // <= 4 columns per thread
  if (dtype == TCD_F32)
    cond_pool_kernel<float><<<n, threads, 0, as_stream(stream)>>>(tokens, null_embed, keep, gamma, beta, (float*)pooled_ln, S, D);
  else if (dtype == TCD_BF16)
    cond_pool_kernel<__nv_bfloat16><<<n, threads, 0, as_stream(stream)>>>(tokens, null_embed, keep, gamma, beta, (__nv_bfloat16*)pooled_ln, S, D);
  else { set_error("tcd_cond_pool: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("cond_pool");
}

extern "C" int tcd_time_cond(int dtype, const float* t_lin, const float* cond_hidden, const float* null_hidden,
                             const uint8_t* keep, float* t_out, void* mish_out, int n, int D, void* stream) {
  TCD_REQUIRE(t_lin && cond_hidden && null_hidden && keep, "tcd_time_cond: null pointer");
  if (n == 0) return TCD_OK;
  const int g = ceil_div((int64_t)n * D, 256);
  if (dtype == TCD_F32)
    time_cond_kernel<float><<<g, 256, 0, as_stream(stream)>>>(t_lin, cond_hidden, null_hidden, keep, t_out, (float*)mish_out, n, D);
  else if (dtype == TCD_BF16)
    time_cond_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>(t_lin, cond_hidden, null_hidden, keep, t_out, (__nv_bfloat16*)mish_out, n, D);
  else { set_error("tcd_time_cond: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("time_cond");
}

extern "C" int tcd_sampler_time_cond(int dtype, const float* t_lin, const float* ch_cond, const float* ch_uncond,
                                     void* mish_out, int steps, int B, int D, void* stream) {
  TCD_REQUIRE(t_lin && ch_cond && ch_uncond && mish_out, "tcd_sampler_time_cond: null pointer");
  const int64_t n = (int64_t)steps * 2 * B * D;
  if (n == 0) return TCD_OK;
  const int g = ceil_div(n, 256);
  if (dtype == TCD_F32)
    sampler_time_cond_kernel<float><<<g, 256, 0, as_stream(stream)>>>(t_lin, ch_cond, ch_uncond, (float*)mish_out, steps, B, D);
  else if (dtype == TCD_BF16)
    sampler_time_cond_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>(t_lin, ch_cond, ch_uncond, (__nv_bfloat16*)mish_out, steps, B, D);
  else { set_error("tcd_sampler_time_cond: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("sampler_time_cond");
}

extern "C" int tcd_time_embed(int dtype, const int64_t* times, const float* table, void* out, int n, int D,
                              int n_timestep, void* stream) {
  TCD_REQUIRE(times && table && out, "tcd_time_embed: null pointer");
  if (n == 0) return TCD_OK;
  const int g = ceil_div((int64_t)n * D, 256);
  if (dtype == TCD_F32)
    time_embed_kernel<float><<<g, 256, 0, as_stream(stream)>>>(times, table, (float*)out, n, D, n_timestep);
  else if (dtype == TCD_BF16)
    time_embed_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>(times, table, (__nv_bfloat16*)out, n, D, n_timestep);
  else { set_error("tcd_time_embed: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("time_embed");
}

extern "C" int tcd_rotary(const float* x, float* out, const float* rot_cos, const float* rot_sin, int64_t rows, int D,
                          int tokens_per_sample, void* stream) {
  TCD_REQUIRE(x && out && rot_cos && rot_sin && D % 2 == 0 && tokens_per_sample > 0, "tcd_rotary: bad arguments");
  if (rows == 0) return TCD_OK;
  rotary_kernel<<<ceil_div(rows * (D / 2), 256), 256, 0, as_stream(stream)>>>(x, out, rot_cos, rot_sin, rows, D, tokens_per_sample);
  return check_launch("rotary");
}

extern "C" int tcd_masked_blend(float* x, const float* value, const float* weight, int B, int S, int dn, int C,
                                void* stream) {
  TCD_REQUIRE(B >= 0 && S > 0 && dn > 0 && C > 0, "tcd_masked_blend: bad shape");
  const int64_t n = (int64_t)B * S * dn * C;
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(x && value && weight, "tcd_masked_blend: null pointer");
  masked_blend_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, value, weight, n, S, dn, C);
  return check_launch("masked_blend");
}

extern "C" int tcd_scatter_rows(int dtype, const void* src, int64_t src_ld, int64_t src_batch_stride, void* dst,
                                int64_t dst_ld, int64_t dst_batch_stride, int64_t dst_row0, int rows, int cols,
                                int samples, void* stream) {
  TCD_REQUIRE(src && dst, "tcd_scatter_rows: null pointer");
  int64_t n = (int64_t)rows * cols * samples;
  if (n == 0) return TCD_OK;
  const int g = ceil_div(n, 256);
  if (dtype == TCD_F32)
    scatter_rows_kernel<float><<<g, 256, 0, as_stream(stream)>>>((const float*)src, src_ld, src_batch_stride, (float*)dst, dst_ld, dst_batch_stride, dst_row0, rows, cols, samples);
  else if (dtype == TCD_BF16)
    scatter_rows_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, src_ld, src_batch_stride, (__nv_bfloat16*)dst, dst_ld, dst_batch_stride, dst_row0, rows, cols, samples);
  else { set_error("tcd_scatter_rows: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("scatter_rows");
}

extern "C" int tcd_convert_pad(int dtype, const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                               int cols, void* stream) {
  TCD_REQUIRE(src && dst && dst_ld >= cols && src_ld >= cols, "tcd_convert_pad: bad arguments");
  if (rows == 0) return TCD_OK;
  const int g = ceil_div(rows * dst_ld, 256);
  if (dtype == TCD_F32)
    convert_pad_kernel<float><<<g, 256, 0, as_stream(stream)>>>(src, src_ld, (float*)dst, dst_ld, rows, cols);
  else if (dtype == TCD_BF16 && cols % 8 == 0 && src_ld % 4 == 0 && dst_ld % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    const int64_t nvec = rows * (dst_ld / 8);
    int gv = ceil_div(nvec, 512);
    if (gv > 148 * 16) gv = 148 * 16;
    convert_bf16_vec8_kernel<<<gv, 256, 0, as_stream(stream)>>>(src, src_ld, (__nv_bfloat16*)dst, dst_ld, rows, cols, (int)(dst_ld / 8));
  } else if (dtype == TCD_BF16)
    convert_pad_kernel<__nv_bfloat16><<<g, 256, 0, as_stream(stream)>>>(src, src_ld, (__nv_bfloat16*)dst, dst_ld, rows, cols);
  else { set_error("tcd_convert_pad: bad dtype %d", dtype); return TCD_ERR_INVALID; }
  return check_launch("convert_pad");
}
