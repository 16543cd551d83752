// Fused diffusion-step kernels (fp32, HBM-bound): CFG blend + x0 clamp + eps + DDIM/DDPM update +
// trajectory in-painting in ONE pass over (n_tokens, 151).  Reference: model/model.py:542-546,
// model/diffusion.py:189-252,385-442,625-651.
//
// Layout: the (n_tokens, 151) tensors are contiguous, so they are streamed as flat float4 vectors
// (16-byte aligned base, 604-byte rows are only 4-byte aligned); the channel index is recovered with
// idx % 151 for the two in-painted channels.  Arithmetic uses explicit round-to-nearest intrinsics in
// the reference's operation order (no FMA contraction) so that, given identical inputs, the update is
// bit-identical to the PyTorch CPU evaluation.
#include "common.cuh"

namespace tcd {

constexpr int kC = 151;

// ---- counter-based Gaussian noise (Philox4x32-10 + Box-Muller) -------------------------------------------------------------
// The reference draws torch.randn_like(x) once per step (model/diffusion.py:246,421).  Here element i of draw `stream` is
// component (i & 3) of the Philox block  counter = {i >> 2 (64 bit), stream, call counter},  key = seed,  with
// {seed, call counter} read from a device-resident pair (the sampler bumps the counter inside its CUDA graph, so a replay
// draws fresh noise; the seed comes from torch's generator).  No noise tensor exists: the 29 MB per step of the c2 sampler
// (a 1.45 GB bank per DDIM-50 call in round 1) are neither written nor read.  torch's own Philox stream cannot be
// reproduced by another implementation; parity tests materialise THIS stream (tcd_philox_normal) into a noise bank.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct RngKey {
  uint2 key;
  uint32_t ctr_lo;
};
__device__ __forceinline__ RngKey rng_key(const uint64_t* __restrict__ state) {
  const uint64_t s = state[0], c = state[1];
  RngKey k;
  k.key = make_uint2((uint32_t)s, (uint32_t)(s >> 32) ^ (uint32_t)(c >> 32) * 0x9E3779B1u);
  k.ctr_lo = (uint32_t)c;
  return k;
}
// four independent N(0,1) values of block `blk` of draw `stream`
__device__ __forceinline__ float4 philox_normal4(const RngKey& k, uint32_t stream, int64_t blk) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)((uint64_t)blk >> 32), stream, k.ctr_lo), k.key);
  const float kU = 5.9604644775390625e-08f;                 // 2^-24: u in (0, 1) strictly (24 random bits + half an ulp)
  const float u0 = fmaf((float)(r.x >> 8), kU, 0.5f * kU), u1 = fmaf((float)(r.y >> 8), kU, 0.5f * kU);
  const float u2 = fmaf((float)(r.z >> 8), kU, 0.5f * kU), u3 = fmaf((float)(r.w >> 8), kU, 0.5f * kU);
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u1, &sa, &ca);
  __sincosf(6.283185307179586f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
__device__ __forceinline__ float philox_normal1(const RngKey& k, uint32_t stream, int64_t i) {
  const float4 v = philox_normal4(k, stream, i >> 2);
  const int c = (int)(i & 3);
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, int64_t n, const uint64_t* __restrict__ state,
                                                            uint32_t stream, int vec) {
  const RngKey k = rng_key(state);
  const int64_t nvec = vec ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride)
    reinterpret_cast<float4*>(out)[v] = philox_normal4(k, stream, v);
  for (int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = philox_normal1(k, stream, t);
}

struct DdimCoef {
  float w, sr, srm1, sa, c, sigma;
  int clip, last;
};

__device__ __forceinline__ float guided(float con, float unc, float w) {
  return __fadd_rn(unc, __fmul_rn(__fsub_rn(con, unc), w));
}

__device__ __forceinline__ float ddim_update(float x, float con, float unc, float nz, const DdimCoef& k,
                                             float& x0) {
  float o = guided(con, unc, k.w);
  x0 = k.clip ? fminf(fmaxf(o, -1.0f), 1.0f) : o;
  if (k.last) return x0;
  float eps = __fdiv_rn(__fsub_rn(__fmul_rn(k.sr, x), x0), k.srm1);
  return __fadd_rn(__fadd_rn(__fmul_rn(x0, k.sa), __fmul_rn(k.c, eps)), __fmul_rn(k.sigma, nz));
}

__device__ __forceinline__ void store_pad(__nv_bfloat16* xpad, int64_t ld, int64_t idx, float v) {
  int64_t tok = idx / kC;
  int ch = (int)(idx - tok * kC);
  xpad[tok * ld + ch] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float traj_override(const float* __restrict__ traj, int64_t idx, float v) {
  int64_t tok = idx / kC;
  int ch = (int)(idx - tok * kC);
  if (ch == 4) return __ldg(traj + tok * 3 + 0);
  if (ch == 5) return __ldg(traj + tok * 3 + 1);
  return v;
}

__global__ void __launch_bounds__(256) cfg_ddim_step_kernel(
    const float* x, const float* __restrict__ con, const float* __restrict__ unc,
    const float* __restrict__ noise, const float* __restrict__ traj, float* x_out,
    float* __restrict__ x0_out, __nv_bfloat16* __restrict__ xpad, int64_t xpad_ld, int64_t n, DdimCoef k,
    int vec, const uint64_t* __restrict__ rng_state, uint32_t rng_stream) {
  // noise == NULL && rng_state != NULL: the step's draw is generated here (philox_normal4), not read
  RngKey rk;
  const bool use_rng = noise == nullptr && rng_state != nullptr && !k.last;
  if (use_rng) rk = rng_key(rng_state);
  // x / x_out carry no __restrict__: the C-ABI allows the update in place (x_out == x).
  // vec == 0: some pointer is only 4-byte aligned (odd B*L slices) -> everything goes through the scalar loop
  const int64_t nvec = vec ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // One 16-byte vector of the four streams -> updated vector (+ x0, + bf16 K-padded copy).  32-bit index arithmetic: the
  // launcher takes this kernel only for n < 2^31 (r02: the 64-bit division per vector and one vector in flight per thread
  // left the launch at 0.60 of HBM at the c2 size; two independent vectors per iteration and 32-bit indices: see DESIGN §5).
  auto one = [&](uint32_t v, float4 xv, float4 cv, float4 uv, float4 nv) {
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, cs[4] = {cv.x, cv.y, cv.z, cv.w};
    const float us[4] = {uv.x, uv.y, uv.z, uv.w}, ns[4] = {nv.x, nv.y, nv.z, nv.w};
    float r[4], z[4];
    const uint32_t i0 = v * 4u;
    uint32_t tok = i0 / (uint32_t)kC;
    int ch = (int)(i0 - tok * (uint32_t)kC);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[j] = ddim_update(xs[j], cs[j], us[j], ns[j], k, z[j]);
      if (traj) {
        if (ch == 4) r[j] = __ldg(traj + (int64_t)tok * 3 + 0);
        else if (ch == 5) r[j] = __ldg(traj + (int64_t)tok * 3 + 1);
      }
      if (xpad) xpad[(int64_t)tok * xpad_ld + ch] = __float2bfloat16_rn(r[j]);
      if (++ch == kC) { ch = 0; ++tok; }
    }
    reinterpret_cast<float4*>(x_out)[v] = make_float4(r[0], r[1], r[2], r[3]);
    if (x0_out) reinterpret_cast<float4*>(x0_out)[v] = make_float4(z[0], z[1], z[2], z[3]);
  };
  auto draw = [&](uint32_t v) {
    if (use_rng) return philox_normal4(rk, rng_stream, (int64_t)v);
    if (!k.last) return __ldg(reinterpret_cast<const float4*>(noise) + v);
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  const uint32_t nv32 = (uint32_t)nvec, st32 = (uint32_t)stride;
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + st32 < nv32; v += 2u * st32) {           // two independent vectors in flight
    const uint32_t w = v + st32;
    const float4 xa = reinterpret_cast<const float4*>(x)[v], xb = reinterpret_cast<const float4*>(x)[w];
    const float4 ca = __ldg(reinterpret_cast<const float4*>(con) + v), cb = __ldg(reinterpret_cast<const float4*>(con) + w);
    const float4 ua = __ldg(reinterpret_cast<const float4*>(unc) + v), ub = __ldg(reinterpret_cast<const float4*>(unc) + w);
    const float4 na = draw(v), nb = draw(w);
    one(v, xa, ca, ua, na);
    one(w, xb, cb, ub, nb);
  }
  if (v < nv32)
    one(v, reinterpret_cast<const float4*>(x)[v], __ldg(reinterpret_cast<const float4*>(con) + v),
        __ldg(reinterpret_cast<const float4*>(unc) + v), draw(v));
  // tail (n % 4 elements), or the whole range when the vector path is disabled
  for (int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    float z;
    float r = ddim_update(x[t], con[t], unc[t], k.last ? 0.f : (use_rng ? philox_normal1(rk, rng_stream, t) : noise[t]), k, z);
    if (traj) r = traj_override(traj, t, r);
    x_out[t] = r;
    if (x0_out) x0_out[t] = z;
    if (xpad) store_pad(xpad, xpad_ld, t, r);
  }
}

struct DdpmCoef {
  float w, c1, c2, nzstd;
  float sr, srm1;           // predict_epsilon: x_recon = sr * x - srm1 * out (model/diffusion.py:181-185)
  int eps;
};

__global__ void __launch_bounds__(256) cfg_ddpm_step_kernel(
    const float* x, const float* __restrict__ con, const float* __restrict__ unc,
    const float* __restrict__ noise, float* x_out, __nv_bfloat16* __restrict__ xpad,
    int64_t xpad_ld, int64_t n, DdpmCoef k, const float* __restrict__ mask, const float* __restrict__ value_q,
    const uint64_t* __restrict__ rng_state, uint32_t rng_stream) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  RngKey rk;
  const bool use_rng = noise == nullptr;
  if (use_rng) rk = rng_key(rng_state);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float o = guided(__ldg(con + i), __ldg(unc + i), k.w);
    const float xi = x[i];
    if (k.eps) o = __fsub_rn(__fmul_rn(k.sr, xi), __fmul_rn(k.srm1, o));
    float x0 = fminf(fmaxf(o, -1.0f), 1.0f);
    float mean = __fadd_rn(__fmul_rn(k.c1, x0), __fmul_rn(k.c2, xi));
    // consecutive threads of a quad recompute the same Philox block (the DDPM kernel is scalar: 151-float rows, any alignment)
    const float nzv = use_rng ? (k.nzstd != 0.f ? philox_normal1(rk, rng_stream, i) : 0.f) : __ldg(noise + i);
    float r = __fadd_rn(mean, __fmul_rn(k.nzstd, nzv));
    if (mask) {
      float m = __ldg(mask + i);
      r = __fadd_rn(__fmul_rn(__ldg(value_q + i), m), __fmul_rn(__fsub_rn(1.0f, m), r));
    }
    x_out[i] = r;
    if (xpad) store_pad(xpad, xpad_ld, i, r);
  }
}

__global__ void __launch_bounds__(256) inpaint_traj_kernel(float* __restrict__ x, const float* __restrict__ traj,
                                                           __nv_bfloat16* __restrict__ xpad, int64_t xpad_ld,
                                                           int64_t n_tokens) {
  // one thread per (token, channel) when a padded copy is wanted, else only channels 4,5
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (xpad) {
    const int64_t n = n_tokens * kC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float v = x[i];
      if (traj) {
        float r = traj_override(traj, i, v);
        if (r != v) x[i] = r;
        v = r;
      }
      store_pad(xpad, xpad_ld, i, v);
    }
  } else {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tokens * 2; t += stride) {
      int64_t tok = t >> 1;
      int j = (int)(t & 1);
      x[tok * kC + 4 + j] = __ldg(traj + tok * 3 + j);
    }
  }
}

__global__ void __launch_bounds__(256) q_sample_kernel(
    const float* __restrict__ x_start, const float* __restrict__ noise, const int64_t* __restrict__ t,
    const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac, float* __restrict__ x_noisy,
    float* __restrict__ target, __nv_bfloat16* __restrict__ xpad, int64_t xpad_ld, int B, int dn, int S,
    int permute, int restore_traj) {
  const int64_t per = (int64_t)dn * S * kC;
  const int64_t n = per * B;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int b = (int)(i / per);
    int64_t r = i - (int64_t)b * per;
    int ch = (int)(r % kC);
    int64_t row = r / kC;  // output row: (s, d) when permuting
    int64_t src = i;
    if (permute) {
      int s = (int)(row / dn), d = (int)(row - (int64_t)s * dn);
      src = (int64_t)b * per + ((int64_t)d * S + s) * kC + ch;
    }
    float xs = __ldg(x_start + src);
    int64_t tb = t[b];
    float v = __fadd_rn(__fmul_rn(__ldg(sqrt_ac + tb), xs), __fmul_rn(__ldg(sqrt_1mac + tb), __ldg(noise + i)));
    if (restore_traj && (ch == 4 || ch == 5)) v = xs;
    x_noisy[i] = v;
    if (target) target[i] = xs;
    if (xpad) store_pad(xpad, xpad_ld, i, v);
  }
}

static int grid_for(int64_t work_items, int block) {
  // HBM-bound streaming: enough CTAs for >= 8 resident per SM on 148 SMs, capped for grid-stride loops
  int64_t g = (work_items + block - 1) / block;
  int64_t cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_cfg_ddim_step(const float* x, const float* out_cond, const float* out_uncond,
                                 const float* noise, const float* traj, float* x_out, float* x0_out,
                                 void* xpad_out, int64_t xpad_ld, int64_t n_tokens, int C, float w,
                                 float sqrt_recip, float sqrt_recipm1, float sqrt_alpha_next, float c,
                                 float sigma, int clip, int last, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_cfg_ddim_step: C must be 151 (model/model.py:553), got %d", C);
  TCD_REQUIRE(n_tokens >= 0, "tcd_cfg_ddim_step: negative size");
  if (n_tokens == 0) return TCD_OK;
  TCD_REQUIRE(x && out_cond && out_uncond && x_out, "tcd_cfg_ddim_step: null pointer");
  TCD_REQUIRE(last || noise, "tcd_cfg_ddim_step: noise required unless last");
  TCD_REQUIRE(!xpad_out || xpad_ld >= C, "tcd_cfg_ddim_step: xpad_ld < C");
  const int64_t n = n_tokens * kC;
  const int vec = ((uintptr_t)x | (uintptr_t)out_cond | (uintptr_t)out_uncond | (uintptr_t)noise | (uintptr_t)x_out |
                   (uintptr_t)x0_out) % 16 == 0 && n < (1LL << 31);
  DdimCoef k{w, sqrt_recip, sqrt_recipm1, sqrt_alpha_next, c, sigma, clip, last};
  cfg_ddim_step_kernel<<<grid_for(vec ? (n >> 2) + 4 : n, 256), 256, 0, as_stream(stream)>>>(
      x, out_cond, out_uncond, noise, traj, x_out, x0_out, (__nv_bfloat16*)xpad_out, xpad_ld, n, k, vec, nullptr, 0u);
  return check_launch("cfg_ddim_step");
}

extern "C" int tcd_cfg_ddim_step_rng(const float* x, const float* out_cond, const float* out_uncond,
                                     const void* rng_state, uint32_t rng_stream, const float* traj, float* x_out,
                                     float* x0_out, void* xpad_out, int64_t xpad_ld, int64_t n_tokens, int C, float w,
                                     float sqrt_recip, float sqrt_recipm1, float sqrt_alpha_next, float c,
                                     float sigma, int clip, int last, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_cfg_ddim_step_rng: C must be 151 (model/model.py:553), got %d", C);
  TCD_REQUIRE(n_tokens >= 0, "tcd_cfg_ddim_step_rng: negative size");
  if (n_tokens == 0) return TCD_OK;
  TCD_REQUIRE(x && out_cond && out_uncond && x_out, "tcd_cfg_ddim_step_rng: null pointer");
  TCD_REQUIRE(last || rng_state, "tcd_cfg_ddim_step_rng: rng_state required unless last");
  TCD_REQUIRE(!xpad_out || xpad_ld >= C, "tcd_cfg_ddim_step_rng: xpad_ld < C");
  const int64_t n = n_tokens * kC;
  const int vec = ((uintptr_t)x | (uintptr_t)out_cond | (uintptr_t)out_uncond | (uintptr_t)x_out | (uintptr_t)x0_out) % 16 == 0 &&
                  n < (1LL << 31);
  DdimCoef k{w, sqrt_recip, sqrt_recipm1, sqrt_alpha_next, c, sigma, clip, last};
  cfg_ddim_step_kernel<<<grid_for(vec ? (n >> 2) + 4 : n, 256), 256, 0, as_stream(stream)>>>(
      x, out_cond, out_uncond, nullptr, traj, x_out, x0_out, (__nv_bfloat16*)xpad_out, xpad_ld, n, k, vec,
      (const uint64_t*)rng_state, rng_stream);
  return check_launch("cfg_ddim_step_rng");
}

extern "C" int tcd_philox_normal(float* out, int64_t n, const void* rng_state, uint32_t rng_stream, void* stream) {
  if (n == 0) return TCD_OK;
  TCD_REQUIRE(out && rng_state && n > 0, "tcd_philox_normal: bad arguments");
  const int vec = (uintptr_t)out % 16 == 0;
  philox_normal_kernel<<<grid_for(vec ? (n >> 2) + 4 : n, 256), 256, 0, as_stream(stream)>>>(out, n, (const uint64_t*)rng_state,
                                                                                          rng_stream, vec);
  return check_launch("philox_normal");
}

extern "C" int tcd_cfg_ddpm_step(const float* x, const float* out_cond, const float* out_uncond,
                                 const float* noise, float* x_out, void* xpad_out, int64_t xpad_ld,
                                 int64_t n_tokens, int C, float w, float coef1, float coef2, float std,
                                 int nonzero, int predict_epsilon, float sqrt_recip, float sqrt_recipm1,
                                 const float* mask, const float* value_q, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_cfg_ddpm_step: C must be 151, got %d", C);
  if (n_tokens == 0) return TCD_OK;
  TCD_REQUIRE(x && out_cond && out_uncond && noise && x_out, "tcd_cfg_ddpm_step: null pointer");
  TCD_REQUIRE((mask == nullptr) == (value_q == nullptr), "tcd_cfg_ddpm_step: mask and value_q go together");
  const int64_t n = n_tokens * kC;
  DdpmCoef k{w, coef1, coef2, nonzero ? std : 0.0f, sqrt_recip, sqrt_recipm1, predict_epsilon};
  cfg_ddpm_step_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      x, out_cond, out_uncond, noise, x_out, (__nv_bfloat16*)xpad_out, xpad_ld, n, k, mask, value_q, nullptr, 0u);
  return check_launch("cfg_ddpm_step");
}

extern "C" int tcd_cfg_ddpm_step_rng(const float* x, const float* out_cond, const float* out_uncond,
                                     const void* rng_state, uint32_t rng_stream, float* x_out, void* xpad_out,
                                     int64_t xpad_ld, int64_t n_tokens, int C, float w, float coef1, float coef2,
                                     float std, int nonzero, int predict_epsilon, float sqrt_recip, float sqrt_recipm1,
                                     const float* mask, const float* value_q, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_cfg_ddpm_step_rng: C must be 151, got %d", C);
  if (n_tokens == 0) return TCD_OK;
  TCD_REQUIRE(x && out_cond && out_uncond && rng_state && x_out, "tcd_cfg_ddpm_step_rng: null pointer");
  TCD_REQUIRE((mask == nullptr) == (value_q == nullptr), "tcd_cfg_ddpm_step_rng: mask and value_q go together");
  const int64_t n = n_tokens * kC;
  DdpmCoef k{w, coef1, coef2, nonzero ? std : 0.0f, sqrt_recip, sqrt_recipm1, predict_epsilon};
  cfg_ddpm_step_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      x, out_cond, out_uncond, nullptr, x_out, (__nv_bfloat16*)xpad_out, xpad_ld, n, k, mask, value_q,
      (const uint64_t*)rng_state, rng_stream);
  return check_launch("cfg_ddpm_step_rng");
}

extern "C" int tcd_inpaint_traj(float* x, const float* traj, void* xpad_out, int64_t xpad_ld, int64_t n_tokens,
                                int C, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_inpaint_traj: C must be 151, got %d", C);
  if (n_tokens == 0) return TCD_OK;
  TCD_REQUIRE(x && (traj || xpad_out), "tcd_inpaint_traj: null pointer");
  int64_t work = xpad_out ? n_tokens * kC : n_tokens * 2;
  inpaint_traj_kernel<<<grid_for(work, 256), 256, 0, as_stream(stream)>>>(x, traj, (__nv_bfloat16*)xpad_out,
                                                                         xpad_ld, n_tokens);
  return check_launch("inpaint_traj");
}

extern "C" int tcd_q_sample(const float* x_start, const float* noise, const int64_t* t, const float* sqrt_ac,
                            const float* sqrt_1mac, float* x_noisy, float* target, void* xpad_out,
                            int64_t xpad_ld, int B, int dn, int S, int C, int permute, int restore_traj,
                            void* stream) {
  TCD_REQUIRE(C == kC, "tcd_q_sample: C must be 151, got %d", C);
  if ((int64_t)B * dn * S == 0) return TCD_OK;
  TCD_REQUIRE(x_start && noise && t && sqrt_ac && sqrt_1mac && x_noisy, "tcd_q_sample: null pointer");
  q_sample_kernel<<<grid_for((int64_t)B * dn * S * kC, 256), 256, 0, as_stream(stream)>>>(
      x_start, noise, t, sqrt_ac, sqrt_1mac, x_noisy, target, (__nv_bfloat16*)xpad_out, xpad_ld, B, dn, S,
      permute, restore_traj);
  return check_launch("q_sample");
}
