// Trajectory front end of the end-to-end test mode: the pieces of TrajDecoder
// (TrajDecoder/model/traj_model.py:125-200) that the denoiser kernels do not already cover, and the Kalman smoother
// that follows it (TrajDecoder/utils/utils_model.py:10-74; the reference runs it as a Python double loop over
// batch x dancers x frames on the host through filterpy 1.4.5).
#include "common.cuh"

namespace tcd {

int attention_f32(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs, const float* V,
                  int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                  float scale, cudaStream_t st);
int attention_f32_hd32(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs, const float* V,
                       int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                       float scale, cudaStream_t st);

constexpr int LH = 64;      // hidden size

// One block per sequence n, one thread per gate row (4H = 256): the thread keeps its W_hh row (64 floats) and W_ih row
// in registers, h lives in shared memory; per time step: gate pre-activations -> sync -> 64 threads update (c, h) -> sync.
__global__ void __launch_bounds__(4 * LH) lstm_layer_kernel(const float* __restrict__ x, int64_t x_ts, int64_t x_ld,
                                                            const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                            const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                            float* __restrict__ out, int64_t out_ts, int64_t out_ld,
                                                            const float* __restrict__ add_table, int T, int I) {
  __shared__ float h[LH];
  __shared__ float xin[LH];
  __shared__ float gates[4 * LH];
  const int n = blockIdx.x, g = threadIdx.x;
  float whh[LH], wih[LH];
#pragma unroll
  for (int k = 0; k < LH; ++k) whh[k] = __ldg(w_hh + g * LH + k);
#pragma unroll
  for (int k = 0; k < LH; ++k) wih[k] = k < I ? __ldg(w_ih + g * I + k) : 0.f;
  const float bias = __ldg(b_ih + g) + __ldg(b_hh + g);
  float c = 0.f;
  const float add = (add_table && g < LH) ? __ldg(add_table + (int64_t)n * LH + g) : 0.f;
  if (g < LH) h[g] = 0.f;
  for (int t = 0; t < T; ++t) {
    if (g < LH) xin[g] = g < I ? __ldg(x + t * x_ts + n * x_ld + g) : 0.f;
    __syncthreads();
    float a = bias;
#pragma unroll
    for (int k = 0; k < LH; ++k) a = fmaf(wih[k], xin[k], a);
#pragma unroll
    for (int k = 0; k < LH; ++k) a = fmaf(whh[k], h[k], a);
    gates[g] = a;
    __syncthreads();
    if (g < LH) {
      const float ig = 1.0f / (1.0f + expf(-gates[g])), fg = 1.0f / (1.0f + expf(-gates[LH + g]));
      const float gg = tanhf(gates[2 * LH + g]), og = 1.0f / (1.0f + expf(-gates[3 * LH + g]));
      c = fg * c + ig * gg;
      const float hv = og * tanhf(c);
      h[g] = hv;
      out[t * out_ts + n * out_ld + g] = hv + add;
    }
    // the next iteration's first __syncthreads orders these writes before the next reads of h
  }
}

// One thread per track; float64 like the numpy/filterpy arithmetic it restates.
__global__ void __launch_bounds__(128) kalman_smooth_kernel(const float* __restrict__ xy, float* __restrict__ out,
                                                            const double* __restrict__ gains, int tracks, int T, double dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tracks) return;
  const float* z = xy + (int64_t)i * T * 2;
  float* o = out + (int64_t)i * T * 2;
  double s0 = z[0], s1 = z[1], s2 = 0.0, s3 = 0.0;                     // (x, y, vx, vy), utils_model.py:57-61
  for (int t = 0; t < T; ++t) {
    // predict: x = F x (constant velocity)
    s0 = s0 + dt * s2;
    s1 = s1 + dt * s3;
    // update: y = z - H x; x = x + K y
    const double y0 = (double)z[2 * t] - s0, y1 = (double)z[2 * t + 1] - s1;
    const double* K = gains + (int64_t)t * 8;
    s0 += K[0] * y0 + K[1] * y1;
    s1 += K[2] * y0 + K[3] * y1;
    s2 += K[4] * y0 + K[5] * y1;
    s3 += K[6] * y0 + K[7] * y1;
    o[2 * t] = (float)s0;
    o[2 * t + 1] = (float)s1;
  }
}

}  // namespace tcd

using namespace tcd;

extern "C" int tcd_lstm_layer(const float* x, int64_t x_ts, int64_t x_ld, const float* w_ih, const float* w_hh,
                              const float* b_ih, const float* b_hh, float* out, int64_t out_ts, int64_t out_ld,
                              const float* add_table, int T, int N, int I, void* stream) {
  TCD_REQUIRE(T >= 0 && N >= 0 && I > 0 && I <= LH, "tcd_lstm_layer: input size must be in [1, 64] (hidden size is 64)");
  if (T == 0 || N == 0) return TCD_OK;
  TCD_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && out, "tcd_lstm_layer: null pointer");
  lstm_layer_kernel<<<N, 4 * LH, 0, as_stream(stream)>>>(x, x_ts, x_ld, w_ih, w_hh, b_ih, b_hh, out, out_ts, out_ld, add_table, T, I);
  return check_launch("lstm_layer");
}

extern "C" int tcd_attention_f32_hd(int head_dim, const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk,
                                    int64_t kbs, const float* V, int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs,
                                    int samples, int heads, int Lq, int Lk, float scale, void* stream) {
  TCD_REQUIRE(samples >= 0 && heads > 0 && Lq >= 0 && Lk > 0, "tcd_attention_f32_hd: bad shape");
  if (samples == 0 || Lq == 0) return TCD_OK;
  TCD_REQUIRE(Q && K && V && O && heads <= 65535 && samples <= 65535, "tcd_attention_f32_hd: bad arguments");
  TCD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && qbs % 4 == 0 && kbs % 4 == 0 && vbs % 4 == 0 &&
              obs % 4 == 0 && ((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0,
              "tcd_attention_f32_hd: 16-byte alignment required");
  if (head_dim == 64)
    return attention_f32(Q, ldq, qbs, K, ldk, kbs, V, ldv, vbs, O, ldo, obs, samples, heads, Lq, Lk, scale, as_stream(stream));
  if (head_dim == 32)
    return attention_f32_hd32(Q, ldq, qbs, K, ldk, kbs, V, ldv, vbs, O, ldo, obs, samples, heads, Lq, Lk, scale, as_stream(stream));
  set_error("tcd_attention_f32_hd: head_dim must be 32 or 64");
  return TCD_ERR_INVALID;
}

extern "C" int tcd_kalman_smooth(const float* xy, float* out, const double* gains, int tracks, int T, double dt, void* stream) {
  TCD_REQUIRE(tracks >= 0 && T >= 0, "tcd_kalman_smooth: bad shape");
  if (tracks == 0 || T == 0) return TCD_OK;
  TCD_REQUIRE(xy && out && gains, "tcd_kalman_smooth: null pointer");
  kalman_smooth_kernel<<<ceil_div(tracks, 128), 128, 0, as_stream(stream)>>>(xy, out, gains, tracks, T, dt);
  return check_launch("kalman_smooth");
}
