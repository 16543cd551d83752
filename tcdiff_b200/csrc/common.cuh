// Shared helpers for the tcdiff_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tcdiff_b200.h"
#include "tuning.cuh"

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

// ---- programmatic dependent launch (csrc/tuning.cuh, TCD_TUNE_PDL) ------------------------------------------------------
// The persistent tensor-core kernels of a denoise step follow each other ~85 times per step.  Launched with the
// programmatic-stream-serialization attribute (launch_pdl below decides per launch), the next kernel's CTAs become resident as soon as every CTA of the running
// one has passed pdl_sync() (or exited) and there is room for them: its prologue (tensor-map prefetch, barrier init,
// tensor-memory allocation) runs beside the running kernel's last tiles, and it then blocks in griddepcontrol.wait until the
// running grid has completed and its writes are visible.  Rules that keep this safe:
//   * every kernel launched through launch_pdl() calls pdl_sync() in ALL threads before it reads or writes anything another
//     kernel of the stream touches (weights and kernel parameters are fine before it);
//   * only kernels whose whole grid is resident at once (one CTA or CTA pair per SM) go through it: a waiting successor
//     holds shared memory, and a predecessor with CTAs still to be scheduled could be starved of it.
__device__ __forceinline__ void pdl_sync() {
#if TCD_TUNE_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// `short_launch`: the launch does at most ~two tiles of work per CTA (the 10-dancer / 1-clip shapes of BASELINE config 4:
// 6 000 rows), where the kernel boundaries are a measurable part of the step (r02: 1.826 -> 1.729 ms per denoise step).  The
// long launches of the batch-64 sampler and of the training step run at the power cap and gain nothing from it
// (profiles/r02_ab.md), so they are launched plainly; pdl_sync() is a no-op for them.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool short_launch,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
#if TCD_TUNE_PDL
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  if (short_launch) {
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
#else
  (void)short_launch;
#endif
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tcd
