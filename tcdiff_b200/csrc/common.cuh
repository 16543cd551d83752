// Shared helpers for the tcdiff_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tcdiff_b200.h"

namespace tcd {

// ---- error plumbing (C-ABI: int return codes + thread-local message) -------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define TCD_REQUIRE(cond, ...)                                                         \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      tcd::set_error(__VA_ARGS__);                                                     \
      return TCD_ERR_INVALID;                                                          \
    }                                                                                  \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- dtype traits -----------------------------------------------------------------------
template <typename T> struct Conv;
template <> struct Conv<float> {
  static __device__ __forceinline__ float to(float v) { return v; }
  static __device__ __forceinline__ float from(float v) { return v; }
};
template <> struct Conv<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 to(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float from(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// ---- activations (exact forms, matching torch CPU semantics) ----------------------------
__device__ __forceinline__ float act_relu(float x) { return x > 0.f ? x : 0.f; }
// F.gelu default = exact erf form (TCDiff.py:85 passes F.gelu)
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch threshold=20
__device__ __forceinline__ float act_mish(float x) { return x * tanhf(softplus_t(x)); }
__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case TCD_ACT_RELU: return act_relu(x);
    case TCD_ACT_GELU: return act_gelu(x);
    case TCD_ACT_MISH: return act_mish(x);
    case TCD_ACT_SILU: return act_silu(x);
    case TCD_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    default: return x;
  }
}

// ---- warp / block reductions ------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace tcd
