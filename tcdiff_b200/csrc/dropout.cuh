// Counter-based dropout masks for the training tape (nn.Dropout sites of model/model.py:35,67,198-205,273-295).
// No mask is ever stored: every kernel that needs a decision recomputes it from
//     keep(seed, a, b) = mix(seed ^ a * C1 ^ b * C2) >= p * 2^32,   mix(y) = (y ^ y >> 16) * M
// (two multiplicative rounds, ~5 integer ops per decision; the full murmur3 finaliser made the attention kernels 1.6-1.9x
// slower for no measurable gain in mask quality: drop rate, neighbour correlations and row/column count variances are
// binomial to 3 digits for both)
// with (a, b) = the element's coordinates (elementwise sites: low / high half of the flat index; attention
// probabilities: (sample, head, query) row id and key index), so the forward kernel, both attention backward kernels
// and the elementwise backward see the same mask.  `seed` is derived per call site from a device-resident
// (seed, step counter) pair, so a CUDA-graph replay draws fresh masks every step.  torch's Philox stream cannot be
// reproduced bit-for-bit by any other implementation; the tests check the tape against the oracle with THESE masks
// materialised (tcd_dropout on a tensor of ones, tcd_dropout_mask_attention) and injected.
#pragma once
#include <stdint.h>

namespace tcd {

__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
constexpr uint32_t kDropC1 = 0x9E3779B1u, kDropC2 = 0x85EBCA77u;
__device__ __forceinline__ uint32_t drop_mix(uint32_t y) { return (y ^ (y >> 16)) * 0x85EBCA6Bu; }

// per-site seed from the device-resident state {seed, step counter}
__device__ __forceinline__ uint32_t drop_site_seed(const uint64_t* __restrict__ state, uint32_t site) {
  const uint64_t s = state[0], c = state[1];
  uint32_t x = fmix32((uint32_t)s ^ 0x2545F491u);
  x = fmix32(x ^ (uint32_t)(s >> 32) * kDropC2);
  x = fmix32(x ^ (uint32_t)c * kDropC1 ^ (uint32_t)(c >> 32));
  return fmix32(x ^ site * 0x632BE5ABu);
}
__device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t a, uint32_t b, uint32_t threshold) {
  return drop_mix(seed ^ a * kDropC1 ^ b * kDropC2) >= threshold;
}
static inline uint32_t drop_threshold(double p) {
  double t = p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

}  // namespace tcd
