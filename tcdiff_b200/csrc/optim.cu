// Optimizer step of the data-parallel training loop: Adan (model/adan.py:33-123) and the EMA blend of the master
// model (model/diffusion.py:61-76, TCDiff.py:242-245) over FLAT fp32 arenas, one launch for the whole model instead of
// ~12 elementwise launches x 446 tensors.  HBM-bound: 7 streams read + 6 written = 52 B per parameter.
//
// Arithmetic follows the reference's op sequence (which products are separate roundings, which are alpha-fused adds)
// so that a step agrees with the torch implementation to the last bit or two; the scalars (1-beta, bias corrections,
// 1+wd*lr) are formed in double on the host exactly like the Python floats they restate and rounded to fp32 once.
#include "common.cuh"

#include <cmath>
#include <math.h>

namespace tcd {

struct AdanScalars {
  float omb1, b1, omb2, b2, omb3, b3;   // (1-beta_i), beta_i
  float cm, cv, cn;                     // bias corrections 1 / (1 - (1-beta_i)^step)
  float lr, eps, denom;                 // denom = 1 + weight_decay * lr
  float grad_scale;                     // 1/world after a SUM all-reduce (1 = untouched)
  float ema_beta, ema_omb;
  int update_moments;                   // 0 on the very first step (adan.py:70: `if step > 0`)
  int has_ema;
};

__device__ __forceinline__ void adan_one(float& p, float g, float& pg, float& m, float& v, float& n, float& e,
                                         const AdanScalars& s) {
  if (s.grad_scale != 1.f) g = __fmul_rn(g, s.grad_scale);
  if (s.update_moments) {
    m = __fmaf_rn(s.b1, g, __fmul_rn(m, s.omb1));                       // m.mul_(1-b1).add_(grad, alpha=b1)
    const float diff = __fsub_rn(g, pg);                                 // grad - prev_grad
    v = __fmaf_rn(s.b2, diff, __fmul_rn(v, s.omb2));                    // v.mul_(1-b2).add_(diff, alpha=b2)
    const float nx = __fadd_rn(g, __fmul_rn(s.omb2, diff));             // grad + (1-b2)*diff
    n = __fmaf_rn(s.b3, __fmul_rn(nx, nx), __fmul_rn(n, s.omb3));       // n.mul_(1-b3).add_(nx**2, alpha=b3)
  }
  const float t = __fadd_rn(__fsqrt_rn(__fmul_rn(n, s.cn)), s.eps);      // (n*correct_n).sqrt().add_(eps)
  const float wss = __fmul_rn(__frcp_rn(t), s.lr);                       // lr / t  ==  t.reciprocal() * lr
  const float upd = __fadd_rn(__fmul_rn(m, s.cm), __fmul_rn(__fmul_rn(s.omb2, v), s.cv));
  p = __fdiv_rn(__fsub_rn(p, __fmul_rn(wss, upd)), s.denom);             // addcmul_(value=-1).div_(denom)
  pg = g;                                                                // prev_grad.copy_(grad)
  if (s.has_ema) e = __fadd_rn(__fmul_rn(e, s.ema_beta), __fmul_rn(s.ema_omb, p));   // old*beta + (1-beta)*new
}

__global__ void __launch_bounds__(256) adan_ema_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                       float* __restrict__ prev_grad, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ n,
                                                       float* __restrict__ ema, int64_t count, AdanScalars s,
                                                       const AdanScalars* __restrict__ s_dev) {
  if (s_dev) s = *s_dev;                    // graph-replayable path: scalars prepared on the device (adan_prepare_kernel)
  const int64_t nvec = count >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 p4 = reinterpret_cast<float4*>(param)[i];
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(grad) + i);
    float4 q4 = reinterpret_cast<float4*>(prev_grad)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    float4 n4 = reinterpret_cast<float4*>(n)[i];
    float4 e4 = s.has_ema ? reinterpret_cast<float4*>(ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adan_one(p4.x, g4.x, q4.x, m4.x, v4.x, n4.x, e4.x, s);
    adan_one(p4.y, g4.y, q4.y, m4.y, v4.y, n4.y, e4.y, s);
    adan_one(p4.z, g4.z, q4.z, m4.z, v4.z, n4.z, e4.z, s);
    adan_one(p4.w, g4.w, q4.w, m4.w, v4.w, n4.w, e4.w, s);
    reinterpret_cast<float4*>(param)[i] = p4;
    reinterpret_cast<float4*>(prev_grad)[i] = q4;
    if (s.update_moments) {
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
      reinterpret_cast<float4*>(n)[i] = n4;
    }
    if (s.has_ema) reinterpret_cast<float4*>(ema)[i] = e4;
  }
  // ragged tail (count not a multiple of 4): one thread each
  const int64_t tail0 = nvec << 2;
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi < count - tail0) {
    const int64_t i = tail0 + gi;
    float e = s.has_ema ? ema[i] : 0.f;
    adan_one(param[i], grad[i], prev_grad[i], m[i], v[i], n[i], e, s);
    if (s.has_ema) ema[i] = e;
  }
}

struct AdanHyper {
  double lr, b1, b2, b3, eps, wd, ema_beta, grad_scale;
  int has_ema;
};

static __host__ __device__ inline AdanScalars make_scalars(const AdanHyper& h, int64_t step) {
  AdanScalars s;
  s.b1 = (float)h.b1; s.omb1 = (float)(1.0 - h.b1);
  s.b2 = (float)h.b2; s.omb2 = (float)(1.0 - h.b2);
  s.b3 = (float)h.b3; s.omb3 = (float)(1.0 - h.b3);
  const double k = (double)(step + 1);                       // adan.py:85 `step += 1` precedes the corrections
  s.cm = (float)(1.0 / (1.0 - pow(1.0 - h.b1, k)));
  s.cv = (float)(1.0 / (1.0 - pow(1.0 - h.b2, k)));
  s.cn = (float)(1.0 / (1.0 - pow(1.0 - h.b3, k)));
  s.lr = (float)h.lr; s.eps = (float)h.eps; s.denom = (float)(1.0 + h.wd * h.lr);
  s.grad_scale = (float)h.grad_scale;
  s.ema_beta = (float)h.ema_beta; s.ema_omb = (float)(1.0 - h.ema_beta);
  s.update_moments = step > 0;
  s.has_ema = h.has_ema;
  return s;
}

// one thread: read and advance the device-resident step counter, derive this step's scalars (CUDA-graph replay safe)
__global__ void adan_prepare_kernel(int64_t* __restrict__ step_dev, AdanHyper h, AdanScalars* __restrict__ out) {
  const int64_t step = *step_dev;
  *out = make_scalars(h, step);
  *step_dev = step + 1;
}

__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ ema, const float* __restrict__ param,
                                                  int64_t count, float beta, float omb) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    ema[i] = __fadd_rn(__fmul_rn(ema[i], beta), __fmul_rn(omb, param[i]));
}

// the same blend for a list of tensors in ONE launch: blockIdx.y = tensor, pointer / size tables in device memory
__global__ void __launch_bounds__(256) ema_multi_kernel(float* const* __restrict__ ema, const float* const* __restrict__ param,
                                                        const int64_t* __restrict__ counts, float beta, float omb) {
  const int t = blockIdx.y;
  float* e = ema[t];
  const float* p = param[t];
  const int64_t n = counts[t];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    e[i] = __fadd_rn(__fmul_rn(e[i], beta), __fmul_rn(omb, p[i]));
}

static int grid_for(int64_t work_items) {
  int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_adan_ema_step(float* param, const float* grad, float* prev_grad, float* exp_avg, float* exp_avg_diff,
                                 float* exp_avg_sq, float* ema, int64_t count, int64_t step, double grad_scale,
                                 double lr, double beta1, double beta2, double beta3, double eps, double weight_decay,
                                 double ema_beta, void* stream) {
  if (count == 0) return TCD_OK;
  TCD_REQUIRE(param && grad && prev_grad && exp_avg && exp_avg_diff && exp_avg_sq, "tcd_adan_ema_step: null pointer");
  TCD_REQUIRE(step >= 0, "tcd_adan_ema_step: step must be the number of updates already applied (>= 0)");
  const uintptr_t bits = (uintptr_t)param | (uintptr_t)grad | (uintptr_t)prev_grad | (uintptr_t)exp_avg |
                         (uintptr_t)exp_avg_diff | (uintptr_t)exp_avg_sq | (uintptr_t)ema;
  TCD_REQUIRE((bits & 15) == 0, "tcd_adan_ema_step: arenas must be 16-byte aligned");
  AdanHyper h{lr, beta1, beta2, beta3, eps, weight_decay, ema_beta, grad_scale, ema != nullptr};
  const AdanScalars s = make_scalars(h, step);
  adan_ema_kernel<<<grid_for((count + 3) / 4), 256, 0, as_stream(stream)>>>(param, grad, prev_grad, exp_avg, exp_avg_diff,
                                                                           exp_avg_sq, ema, count, s, nullptr);
  return check_launch("adan_ema_step");
}

extern "C" int tcd_adan_ema_step_device(float* param, const float* grad, float* prev_grad, float* exp_avg, float* exp_avg_diff,
                                        float* exp_avg_sq, float* ema, int64_t count, int64_t* step_device,
                                        void* scalars_workspace, double grad_scale, double lr, double beta1, double beta2,
                                        double beta3, double eps, double weight_decay, double ema_beta, void* stream) {
  if (count == 0) return TCD_OK;
  TCD_REQUIRE(param && grad && prev_grad && exp_avg && exp_avg_diff && exp_avg_sq && step_device && scalars_workspace,
              "tcd_adan_ema_step_device: null pointer");
  const uintptr_t bits = (uintptr_t)param | (uintptr_t)grad | (uintptr_t)prev_grad | (uintptr_t)exp_avg |
                         (uintptr_t)exp_avg_diff | (uintptr_t)exp_avg_sq | (uintptr_t)ema | (uintptr_t)scalars_workspace;
  TCD_REQUIRE((bits & 15) == 0 && (uintptr_t)step_device % 8 == 0, "tcd_adan_ema_step_device: arenas must be 16-byte aligned");
  static_assert(sizeof(AdanScalars) <= 128, "scalars workspace is 128 bytes");
  AdanHyper h{lr, beta1, beta2, beta3, eps, weight_decay, ema_beta, grad_scale, ema != nullptr};
  AdanScalars* sp = reinterpret_cast<AdanScalars*>(scalars_workspace);
  adan_prepare_kernel<<<1, 1, 0, as_stream(stream)>>>(step_device, h, sp);
  int rc = check_launch("adan_prepare");
  if (rc) return rc;
  adan_ema_kernel<<<grid_for((count + 3) / 4), 256, 0, as_stream(stream)>>>(param, grad, prev_grad, exp_avg, exp_avg_diff,
                                                                           exp_avg_sq, ema, count, AdanScalars{}, sp);
  return check_launch("adan_ema_step_device");
}

extern "C" int tcd_ema_update(float* ema, const float* param, int64_t count, double beta, void* stream) {
  if (count == 0) return TCD_OK;
  TCD_REQUIRE(ema && param, "tcd_ema_update: null pointer");
  ema_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(ema, param, count, (float)beta, (float)(1.0 - beta));
  return check_launch("ema_update");
}

extern "C" int tcd_ema_update_multi(const void* ema_ptrs, const void* param_ptrs, const int64_t* counts, int n_tensors,
                                    int64_t max_count, double beta, void* stream) {
  if (n_tensors == 0) return TCD_OK;
  TCD_REQUIRE(ema_ptrs && param_ptrs && counts && n_tensors > 0 && n_tensors <= 65535 && max_count >= 0,
              "tcd_ema_update_multi: bad arguments");
  int64_t bx = (max_count + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  ema_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_tensors), 256, 0, as_stream(stream)>>>(
      (float* const*)ema_ptrs, (const float* const*)param_ptrs, counts, (float)beta, (float)(1.0 - beta));
  return check_launch("ema_update_multi");
}
