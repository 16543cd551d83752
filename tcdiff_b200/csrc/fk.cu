// Kinematics and training-loss kernels (fp32).
//   tcd_ax_from_6v  — faithful restatement of pytorch3d 0.7.1 rotation_6d_to_matrix + matrix_to_axis_angle
//                     (dataset/quaternion.py:28-32)
//   tcd_smpl_fk     — faithful quaternion-chain FK of SMPLSkeleton.forward (vis.py:358-406)
//   tcd_motion_fk   — fused 6D -> joints via a direct rotation-matrix chain (model/diffusion.py:692-708)
//   tcd_loss_forward— the four p_losses terms (model/diffusion.py:664-741) in one pass over the
//                     prediction/target rows, with a deterministic two-stage reduction.
#include <utility>

#include "common.cuh"

namespace tcd {

constexpr int kC = 151;
constexpr int kJ = 24;

// vis.py:48-73
__constant__ int c_parent[kJ] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
// vis.py:76-101
__constant__ float c_off[kJ][3] = {
    {0.0f, 0.0f, 0.0f},
    {0.05858135f, -0.08228004f, -0.01766408f},
    {-0.06030973f, -0.09051332f, -0.01354254f},
    {0.00443945f, 0.12440352f, -0.03838522f},
    {0.04345142f, -0.38646945f, 0.008037f},
    {-0.04325663f, -0.38368791f, -0.00484304f},
    {0.00448844f, 0.1379564f, 0.02682033f},
    {-0.01479032f, -0.42687458f, -0.037428f},
    {0.01905555f, -0.4200455f, -0.03456167f},
    {-0.00226458f, 0.05603239f, 0.00285505f},
    {0.04105436f, -0.06028581f, 0.12204243f},
    {-0.03483987f, -0.06210566f, 0.13032329f},
    {-0.0133902f, 0.21163553f, -0.03346758f},
    {0.07170245f, 0.11399969f, -0.01889817f},
    {-0.08295366f, 0.11247234f, -0.02370739f},
    {0.01011321f, 0.08893734f, 0.05040987f},
    {0.12292141f, 0.04520509f, -0.019046f},
    {-0.11322832f, 0.04685326f, -0.00847207f},
    {0.2553319f, -0.01564902f, -0.02294649f},
    {-0.26012748f, -0.01436928f, -0.03126873f},
    {0.26570925f, 0.01269811f, -0.00737473f},
    {-0.26910836f, 0.00679372f, -0.00602676f},
    {0.08669055f, -0.01063603f, -0.01559429f},
    {-0.0887537f, -0.00865157f, -0.01010708f}};

// compile-time copies for fully unrolled chains
#define TCD_PARENTS {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21}
#define TCD_HAS_CHILD {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0}

// ---------------------------------------------------------------------------------------------
// pytorch3d-faithful pieces
// ---------------------------------------------------------------------------------------------
struct Quat { float w, x, y, z; };

__device__ __forceinline__ void rot6d_to_rows(const float* a, float* b1, float* b2, float* b3) {
  // F.normalize: v / max(||v||, 1e-12)
  float n1 = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  float i1 = 1.0f / fmaxf(n1, 1e-12f);
  b1[0] = a[0] * i1; b1[1] = a[1] * i1; b1[2] = a[2] * i1;
  float d = b1[0] * a[3] + b1[1] * a[4] + b1[2] * a[5];
  float u0 = a[3] - d * b1[0], u1 = a[4] - d * b1[1], u2 = a[5] - d * b1[2];
  float n2 = sqrtf(u0 * u0 + u1 * u1 + u2 * u2);
  float i2 = 1.0f / fmaxf(n2, 1e-12f);
  b2[0] = u0 * i2; b2[1] = u1 * i2; b2[2] = u2 * i2;
  b3[0] = b1[1] * b2[2] - b1[2] * b2[1];
  b3[1] = b1[2] * b2[0] - b1[0] * b2[2];
  b3[2] = b1[0] * b2[1] - b1[1] * b2[0];
}

__device__ __forceinline__ Quat rows_to_quat(const float* r0, const float* r1, const float* r2) {
  // matrix_to_quaternion @0.7.1: 4 candidates, argmax of q_abs (first on ties), / (2*max(q_abs,0.1))
  float m00 = r0[0], m01 = r0[1], m02 = r0[2], m10 = r1[0], m11 = r1[1], m12 = r1[2];
  float m20 = r2[0], m21 = r2[1], m22 = r2[2];
  float t0 = 1.0f + m00 + m11 + m22, t1 = 1.0f + m00 - m11 - m22;
  float t2 = 1.0f - m00 + m11 - m22, t3 = 1.0f - m00 - m11 + m22;
  float q0 = t0 > 0.f ? sqrtf(t0) : 0.f, q1 = t1 > 0.f ? sqrtf(t1) : 0.f;
  float q2 = t2 > 0.f ? sqrtf(t2) : 0.f, q3 = t3 > 0.f ? sqrtf(t3) : 0.f;
  int best = 0; float bv = q0;
  if (q1 > bv) { bv = q1; best = 1; }
  if (q2 > bv) { bv = q2; best = 2; }
  if (q3 > bv) { bv = q3; best = 3; }
  float den = 2.0f * fmaxf(bv, 0.1f);
  Quat q;
  if (best == 0)      { q.w = q0 * q0;   q.x = m21 - m12; q.y = m02 - m20; q.z = m10 - m01; }
  else if (best == 1) { q.w = m21 - m12; q.x = q1 * q1;   q.y = m10 + m01; q.z = m02 + m20; }
  else if (best == 2) { q.w = m02 - m20; q.x = m10 + m01; q.y = q2 * q2;   q.z = m12 + m21; }
  else                { q.w = m10 - m01; q.x = m20 + m02; q.y = m21 + m12; q.z = q3 * q3; }
  q.w /= den; q.x /= den; q.y /= den; q.z /= den;
  return q;
}

__device__ __forceinline__ void quat_to_aa(const Quat& q, float* aa) {
  float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  float half = atan2f(n, q.w);
  float ang = 2.0f * half;
  float k = fabsf(ang) < 1e-6f ? 0.5f - (ang * ang) / 48.0f : sinf(half) / ang;
  aa[0] = q.x / k; aa[1] = q.y / k; aa[2] = q.z / k;
}

__device__ __forceinline__ Quat aa_to_quat(const float* aa) {
  float ang = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  float half = 0.5f * ang;
  float k = fabsf(ang) < 1e-6f ? 0.5f - (ang * ang) / 48.0f : sinf(half) / ang;
  Quat q{cosf(half), aa[0] * k, aa[1] * k, aa[2] * k};
  return q;
}

__device__ __forceinline__ Quat qmul_raw(const Quat& a, const Quat& b) {
  Quat o;
  o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return o;
}

__global__ void __launch_bounds__(256) ax_from_6v_kernel(const float* __restrict__ d6, float* __restrict__ aa,
                                                         int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = __ldg(d6 + i * 6 + k);
  float b1[3], b2[3], b3[3], out[3];
  rot6d_to_rows(a, b1, b2, b3);
  quat_to_aa(rows_to_quat(b1, b2, b3), out);
  aa[i * 3 + 0] = out[0]; aa[i * 3 + 1] = out[1]; aa[i * 3 + 2] = out[2];
}

// One thread per skeleton; world quaternions live in a small per-thread array (dynamic parent index
// through constant memory; this is the API-parity kernel, the fused loss path uses the matrix chain).
__global__ void __launch_bounds__(128) smpl_fk_kernel(const float* __restrict__ aa, const float* __restrict__ root,
                                                      float* __restrict__ pos, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;
  Quat rw[kJ];
  float pw[kJ][3];
  const float* a = aa + i * (kJ * 3);
  float* o = pos + i * (kJ * 3);
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    float v[3] = {__ldg(a + j * 3), __ldg(a + j * 3 + 1), __ldg(a + j * 3 + 2)};
    Quat ql = aa_to_quat(v);
    constexpr int parents[kJ] = TCD_PARENTS;
    const int p = parents[j];
    if (p < 0) {
      rw[j] = ql;
      pw[j][0] = __ldg(root + i * 3); pw[j][1] = __ldg(root + i * 3 + 1); pw[j][2] = __ldg(root + i * 3 + 2);
    } else {
      // quaternion_apply(q, off) = (q (x) (0,off) (x) conj(q)).xyz, two raw products
      Quat pq{0.f, c_off[j][0], c_off[j][1], c_off[j][2]};
      Quat conj{rw[p].w, -rw[p].x, -rw[p].y, -rw[p].z};
      Quat r = qmul_raw(qmul_raw(rw[p], pq), conj);
      pw[j][0] = r.x + pw[p][0]; pw[j][1] = r.y + pw[p][1]; pw[j][2] = r.z + pw[p][2];
      if (has_child[j]) {
        Quat m = qmul_raw(rw[p], ql);
        if (m.w < 0.f) { m.w = -m.w; m.x = -m.x; m.y = -m.y; m.z = -m.z; }  // standardize_quaternion
        rw[j] = m;
      }
    }
    o[j * 3] = pw[j][0]; o[j * 3 + 1] = pw[j][1]; o[j * 3 + 2] = pw[j][2];
  }
}

// ---------------------------------------------------------------------------------------------
// Direct rotation-matrix chain from a motion row [contact4 | root3 | 24 x rot6d] (row may live in
// shared or global memory).  out: 24 x 3 positions with stride `os` floats between joints' components
// packed as out[j*3+c].  Mathematically identical to 6D -> R -> quat -> axis-angle -> quat -> FK.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fk_from_row(const float* row, float* out) {
  constexpr int parents[kJ] = TCD_PARENTS;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;
  float R[kJ][9];
  float P[kJ][3];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    float a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = row[7 + j * 6 + k];
    float L[9];
    rot6d_to_rows(a, L, L + 3, L + 6);
    const int p = parents[j];
    if (p < 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[j][k] = L[k];
      P[j][0] = row[4]; P[j][1] = row[5]; P[j][2] = row[6];
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
        P[j][r] = R[p][r * 3] * c_off[j][0] + R[p][r * 3 + 1] * c_off[j][1] + R[p][r * 3 + 2] * c_off[j][2] + P[p][r];
      if (has_child[j]) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            R[j][r * 3 + c] = R[p][r * 3] * L[c] + R[p][r * 3 + 1] * L[3 + c] + R[p][r * 3 + 2] * L[6 + c];
      }
    }
    out[j * 3] = P[j][0]; out[j * 3 + 1] = P[j][1]; out[j * 3 + 2] = P[j][2];
  }
}

constexpr int kPosStride = 73;  // 72 floats per skeleton + 1 pad: odd stride => conflict-free row access

__global__ void __launch_bounds__(128) motion_fk_kernel(const float* __restrict__ motion, float* __restrict__ pos,
                                                        int64_t n) {
  // 32 rows per block staged through shared memory so that both the 604-byte row reads and the
  // 288-byte position writes are coalesced.
  __shared__ float s_row[32 * kC];
  __shared__ float s_pos[32 * kPosStride];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int rows = (int)((n - r0) < 32 ? (n - r0) : 32);
  for (int i = threadIdx.x; i < rows * kC; i += blockDim.x) s_row[i] = __ldg(motion + r0 * kC + i);
  __syncthreads();
  if (threadIdx.x < rows) fk_from_row(s_row + threadIdx.x * kC, s_pos + threadIdx.x * kPosStride);
  __syncthreads();
  for (int i = threadIdx.x; i < rows * 72; i += blockDim.x) {
    int r = i / 72, c = i - r * 72;
    pos[r0 * 72 + i] = s_pos[r * kPosStride + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Fused loss forward.  One WARP owns 32 consecutive rows (s,d) of one sample; consecutive tiles overlap by dn rows
// (the same dancers one frame later), so the last dn lanes of a tile only provide the "next frame" for the frame
// differences and their own terms are counted by the next tile.  The 151-float rows are streamed through a small
// per-warp shared-memory buffer in column chunks (header 7, then 5 joints = 30 columns at a time):
//   * staging: lane = column, one coalesced 120-byte segment per row and tensor;
//   * reconstruction / velocity terms: lane = column, rows walked in shared memory;
//   * kinematics: lane = row (stride-31 rows are bank-conflict free); the prediction and target chains advance
//     together joint by joint with their live world matrices in registers, and the root-relative joint error is
//     accumulated on the fly, so no joint positions are stored except the 4 foot joints (smem exchange).
// r01 profile of the previous block-tile version (thread-per-skeleton on rows staged whole in smem): 0.59 TB/s,
// warps active 18 %, issue 26 % — ~6 working warps per SM; traffic == algorithmic bytes.
// ---------------------------------------------------------------------------------------------
constexpr int kLossWarps = 4;
constexpr int kLossBwdWarps = 4;
constexpr int kChunkStride = 31;                         // <= 31 columns per chunk; odd stride: conflict-free both ways
constexpr int kWarpSmemFloats = 2 * 32 * kChunkStride + 32 * 12;
constexpr int kLossStageRows = 16;                       // rows x 2 tensors of 4-byte loads in flight per lane while staging

__device__ __forceinline__ void rot6d_to_rows_fast(const float* a, float* L) {
  // same arithmetic as rotation_6d_to_matrix (F.normalize eps 1e-12 <=> clamp of the squared norm at 1e-24)
  const float i1 = rsqrtf(fmaxf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2], 1e-24f));
  L[0] = a[0] * i1; L[1] = a[1] * i1; L[2] = a[2] * i1;
  const float d = L[0] * a[3] + L[1] * a[4] + L[2] * a[5];
  const float u0 = a[3] - d * L[0], u1 = a[4] - d * L[1], u2 = a[5] - d * L[2];
  const float i2 = rsqrtf(fmaxf(u0 * u0 + u1 * u1 + u2 * u2, 1e-24f));
  L[3] = u0 * i2; L[4] = u1 * i2; L[5] = u2 * i2;
  L[6] = L[1] * L[5] - L[2] * L[4];
  L[7] = L[2] * L[3] - L[0] * L[5];
  L[8] = L[0] * L[4] - L[1] * L[3];
}

// element loss of the four terms and its derivative without the constant: F.mse_loss (d^2, 2 d) or F.l1_loss (|d|, sign d),
// model/diffusion.py:172 (the reference constructor defaults to l1; TCDiff.py:90-102 passes l2)
template <bool L1>
__device__ __forceinline__ float loss_el(float d) { return L1 ? fabsf(d) : d * d; }
template <bool L1>
__device__ __forceinline__ float loss_del(float d) { return L1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : d; }

// Both chains of a row advance together on the packed fp32 pipe: every value is a float2 (x = prediction, y = target), every
// operation of the chain is one fma.rn.f32x2 / mul.rn.f32x2 for the two skeletons (the scalar version issued 18.2 M warp
// instructions per batch-128 call, 55 % of them these chains: instruction-issue bound at 0.44 of that roofline,
// profiles/r01_loss_kernel.md).  Root position zero in both chains: the FK term is root-relative (model/diffusion.py:712-713)
// and only the four foot joints need the absolute position back.
using f2 = float2;
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 rsqrt_clamped2(f2 n) {     // F.normalize eps 1e-12 <=> clamp of the squared norm at 1e-24
  f2 r;                                                    // (the clamped argument is never subnormal: .ftz drops rsqrtf's rescaling)
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(fmaxf(n.x, 1e-24f)));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(fmaxf(n.y, 1e-24f)));
  return r;
}
__device__ __forceinline__ void rot6d_to_rows2(const f2* a, f2* L) {
  const f2 i1 = rsqrt_clamped2(fma2(a[2], a[2], fma2(a[1], a[1], mul2(a[0], a[0]))));
  L[0] = mul2(a[0], i1); L[1] = mul2(a[1], i1); L[2] = mul2(a[2], i1);
  const f2 nd = neg2(fma2(L[2], a[5], fma2(L[1], a[4], mul2(L[0], a[3]))));
  const f2 u0 = fma2(nd, L[0], a[3]), u1 = fma2(nd, L[1], a[4]), u2 = fma2(nd, L[2], a[5]);
  const f2 i2 = rsqrt_clamped2(fma2(u2, u2, fma2(u1, u1, mul2(u0, u0))));
  L[3] = mul2(u0, i2); L[4] = mul2(u1, i2); L[5] = mul2(u2, i2);
  L[6] = fma2(L[1], L[5], neg2(mul2(L[2], L[4])));
  L[7] = fma2(L[2], L[3], neg2(mul2(L[0], L[5])));
  L[8] = fma2(L[0], L[4], neg2(mul2(L[1], L[3])));
}

// joint offsets as (c, c) pairs for the packed pipe
__constant__ float2 c_off2[kJ][3] = {
#define TCD_O2(a, b, c) {{a, a}, {b, b}, {c, c}}
    TCD_O2(0.0f, 0.0f, 0.0f),
    TCD_O2(0.05858135f, -0.08228004f, -0.01766408f),
    TCD_O2(-0.06030973f, -0.09051332f, -0.01354254f),
    TCD_O2(0.00443945f, 0.12440352f, -0.03838522f),
    TCD_O2(0.04345142f, -0.38646945f, 0.008037f),
    TCD_O2(-0.04325663f, -0.38368791f, -0.00484304f),
    TCD_O2(0.00448844f, 0.1379564f, 0.02682033f),
    TCD_O2(-0.01479032f, -0.42687458f, -0.037428f),
    TCD_O2(0.01905555f, -0.4200455f, -0.03456167f),
    TCD_O2(-0.00226458f, 0.05603239f, 0.00285505f),
    TCD_O2(0.04105436f, -0.06028581f, 0.12204243f),
    TCD_O2(-0.03483987f, -0.06210566f, 0.13032329f),
    TCD_O2(-0.0133902f, 0.21163553f, -0.03346758f),
    TCD_O2(0.07170245f, 0.11399969f, -0.01889817f),
    TCD_O2(-0.08295366f, 0.11247234f, -0.02370739f),
    TCD_O2(0.01011321f, 0.08893734f, 0.05040987f),
    TCD_O2(0.12292141f, 0.04520509f, -0.019046f),
    TCD_O2(-0.11322832f, 0.04685326f, -0.00847207f),
    TCD_O2(0.2553319f, -0.01564902f, -0.02294649f),
    TCD_O2(-0.26012748f, -0.01436928f, -0.03126873f),
    TCD_O2(0.26570925f, 0.01269811f, -0.00737473f),
    TCD_O2(-0.26910836f, 0.00679372f, -0.00602676f),
    TCD_O2(0.08669055f, -0.01063603f, -0.01559429f),
    TCD_O2(-0.0887537f, -0.00865157f, -0.01010708f)
#undef TCD_O2
};

// one joint of both chains: world rotation / root-relative position from the parent's (compile-time tree)
template <int J>
__device__ __forceinline__ void chain_joint2(const f2* a6, f2 (&R)[kJ][9], f2 (&P)[kJ][3]) {
  constexpr int parents[kJ] = TCD_PARENTS;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;
  constexpr int p = parents[J];
  if constexpr (has_child[J]) {
    f2 L[9];
    rot6d_to_rows2(a6, L);
    if constexpr (p < 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[J][k] = L[k];
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          R[J][r * 3 + c] = fma2(R[p][r * 3 + 2], L[6 + c], fma2(R[p][r * 3 + 1], L[3 + c], mul2(R[p][r * 3], L[c])));
    }
  }
  if constexpr (p >= 0) {
    const f2 c0 = c_off2[J][0], c1 = c_off2[J][1], c2 = c_off2[J][2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if constexpr (p == 0) P[J][r] = fma2(R[p][r * 3 + 2], c2, fma2(R[p][r * 3 + 1], c1, mul2(R[p][r * 3], c0)));  // root at the origin
      else                  P[J][r] = fma2(R[p][r * 3 + 2], c2, fma2(R[p][r * 3 + 1], c1, fma2(R[p][r * 3], c0, P[p][r])));
    }
  }
}

// joints [J, JEND) of both chains for this lane's row; rm / rt point at the lane's chunk values (6 per joint from J0).
// The reconstruction and velocity terms of the same 6 columns are taken here too, lane = row: d = prediction - target
// once per element, the next frame's d by one shuffle ((m' - m) - (t' - t) == d' - d up to fp32 rounding).
template <bool L1, int J, int JEND, int J0>
__device__ __forceinline__ void chain_range2(const float* rm, const float* rt, f2 (&R)[kJ][9], f2 (&P)[kJ][3], const float* root,
                                             int dn, float& rec, float& vel, float& fk, float* feet) {
  f2 a[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    a[k] = make_float2(rm[(J - J0) * 6 + k], rt[(J - J0) * 6 + k]);
    const float d = a[k].x - a[k].y;
    rec += loss_el<L1>(d);
    vel += loss_el<L1>(__shfl_down_sync(0xffffffffu, d, dn) - d);
  }
  chain_joint2<J>(a, R, P);
  if constexpr (J > 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) fk += loss_el<L1>(P[J][c].x - P[J][c].y);
  }
  if constexpr (J == 7 || J == 8 || J == 10 || J == 11) {
    constexpr int f = J == 7 ? 0 : (J == 8 ? 1 : (J == 10 ? 2 : 3));
#pragma unroll
    for (int c = 0; c < 3; ++c) feet[f * 3 + c] = P[J][c].x + root[c];
  }
  if constexpr (J + 1 < JEND) chain_range2<L1, J + 1, JEND, J0>(rm, rt, R, P, root, dn, rec, vel, fk, feet);
}

template <bool L1>
__global__ void __launch_bounds__(kLossWarps * 32, 4) loss_forward_kernel(
    const float* __restrict__ model_out, const float* __restrict__ target, float* __restrict__ partial, int S, int dn,
    int tiles_per_sample, int total_tiles) {
  __shared__ float smem[kLossWarps * kWarpSmemFloats];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wt = blockIdx.x * kLossWarps + warp;
  if (wt >= total_tiles) return;                           // whole warps only; no block-level barrier below
  float* sm = smem + warp * kWarpSmemFloats;
  float* st = sm + 32 * kChunkStride;
  float* sfeet = st + 32 * kChunkStride;
  const int rows_per_sample = S * dn;
  const int b = wt / tiles_per_sample, tile = wt - b * tiles_per_sample;
  const int adv = 32 - dn;                                 // rows this tile is responsible for
  const int r0 = tile * adv;
  const int nrows = min(32, rows_per_sample - r0);         // rows of the sample present in this tile
  const int nmain = min(adv, nrows);
  const float* gm = model_out + ((int64_t)b * rows_per_sample + r0) * kC;
  const float* gt = target + ((int64_t)b * rows_per_sample + r0) * kC;
  const bool pair_ok = lane < nmain && r0 + lane + dn < rows_per_sample;      // this row has a next frame

  float rec = 0.f, vel = 0.f, fk = 0.f, foot = 0.f;
  f2 R[kJ][9], P[kJ][3];
  float contact[4], root[3], feet[12];

  // staging of one column chunk, lane = column: 2 x 16 independent 4-byte loads in flight per lane (rows are only 4-byte
  // aligned), then the stores; absent rows of a ragged last tile are staged as zeros.  Measured and dropped
  // (profiles/r02_loss_kernel.md): 4-byte cp.async for all 64 segments of a chunk (one round trip, but the copies queue
  // in the memory-input path: mio / lg throttle stalls), the same double-buffered (slower again), 8 or 32 rows in flight,
  // three blocks per SM at 145 / 167 registers, an up-front L2 prefetch of the tile (all within +-5 %, none better).
  auto stage = [&](int col0, int nc) {
    const bool col_ok = lane < nc;
    const float* pm = gm + col0 + lane;
    const float* pt = gt + col0 + lane;
    constexpr int NR = kLossStageRows;
#pragma unroll 1
    for (int rb = 0; rb < 32; rb += NR) {
      float vm[NR], vt[NR];
      if (nrows == 32) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          vm[i] = col_ok ? __ldg(pm + (rb + i) * kC) : 0.f;
          vt[i] = col_ok ? __ldg(pt + (rb + i) * kC) : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const bool ok = col_ok && (rb + i < nrows);
          vm[i] = ok ? __ldg(pm + (rb + i) * kC) : 0.f;
          vt[i] = ok ? __ldg(pt + (rb + i) * kC) : 0.f;
        }
      }
      if (col_ok) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          sm[(rb + i) * kChunkStride + lane] = vm[i];
          st[(rb + i) * kChunkStride + lane] = vt[i];
        }
      }
    }
    __syncwarp();
  };
  const float* rm = sm + lane * kChunkStride;              // lane = row from here on (stride 31: conflict-free)
  const float* rt = st + lane * kChunkStride;

  // ---- chunk 0: contact(4) + root(3) + joints 0..3; the velocity term covers channels 4..150 (model/diffusion.py:672-681)
  stage(0, 31);
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const float m = rm[k], d = m - rt[k];
    rec += loss_el<L1>(d);
    if (k < 4) contact[k] = m;
    else { root[k - 4] = m; vel += loss_el<L1>(__shfl_down_sync(0xffffffffu, d, dn) - d); }
  }
  chain_range2<L1, 0, 4, 0>(rm + 7, rt + 7, R, P, root, dn, rec, vel, fk, feet);       __syncwarp();
  // ---- joints 4..23, 5 per chunk
  stage(31, 30);      chain_range2<L1, 4, 9, 4>(rm, rt, R, P, root, dn, rec, vel, fk, feet);     __syncwarp();
  stage(61, 30);      chain_range2<L1, 9, 14, 9>(rm, rt, R, P, root, dn, rec, vel, fk, feet);    __syncwarp();
  stage(91, 30);      chain_range2<L1, 14, 19, 14>(rm, rt, R, P, root, dn, rec, vel, fk, feet);  __syncwarp();
  stage(121, 30);     chain_range2<L1, 19, 24, 19>(rm, rt, R, P, root, dn, rec, vel, fk, feet);  __syncwarp();
  if (lane >= nmain) { rec = 0.f; fk = 0.f; }              // halo / absent rows are counted by the next tile
  if (!pair_ok) vel = 0.f;
  // ---- foot skate: velocity of joints 7,8,10,11 where the predicted contact > 0.95 (:720-733)
#pragma unroll
  for (int k = 0; k < 12; ++k) sfeet[lane * 12 + k] = feet[k];
  __syncwarp();
  if (pair_ok) {
#pragma unroll
    for (int f = 0; f < 4; ++f)
      if (contact[f] > 0.95f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = sfeet[(lane + dn) * 12 + f * 3 + c] - feet[f * 3 + c];
          foot += loss_el<L1>(v);
        }
      }
  }
  rec = warp_sum(rec); vel = warp_sum(vel); fk = warp_sum(fk); foot = warp_sum(foot);
  if (lane == 0) {
    float* p = partial + (int64_t)wt * 4;
    p[0] = rec; p[1] = vel; p[2] = fk; p[3] = foot;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused loss backward: d total / d model_out for the four terms of tcd_loss_forward.
// Warp per 32 rows again, but with a halo of dn rows on BOTH sides (a row's foot-skate gradient couples it to the same
// dancer one frame earlier and later), lanes [dn, 32-dn) own their rows.  Per lane: forward chains of prediction and
// target (local rotations L and world rotations R of the prediction kept for the reverse sweep), joint-position
// gradients from the FK and foot terms, reverse sweep over the kinematic tree, Gram-Schmidt backward to the 6-D
// inputs; the reconstruction / velocity gradients are added column-wise while each chunk is written out coalesced.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot6d_backward(const float* a, const float* gL, float* ga) {
  // forward (recomputed): b1 = a1/|a1|, u = a2 - (b1.a2) b1, b2 = u/|u|, b3 = b1 x b2
  const float n1sq = fmaxf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2], 1e-24f);
  const float i1 = rsqrtf(n1sq);
  const float b1[3] = {a[0] * i1, a[1] * i1, a[2] * i1};
  const float d = b1[0] * a[3] + b1[1] * a[4] + b1[2] * a[5];
  const float u[3] = {a[3] - d * b1[0], a[4] - d * b1[1], a[5] - d * b1[2]};
  const float i2 = rsqrtf(fmaxf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2], 1e-24f));
  const float b2[3] = {u[0] * i2, u[1] * i2, u[2] * i2};
  float g1[3] = {gL[0], gL[1], gL[2]}, g2[3] = {gL[3], gL[4], gL[5]};
  const float g3[3] = {gL[6], gL[7], gL[8]};
  // b3 = b1 x b2
  g1[0] += b2[1] * g3[2] - b2[2] * g3[1]; g1[1] += b2[2] * g3[0] - b2[0] * g3[2]; g1[2] += b2[0] * g3[1] - b2[1] * g3[0];
  g2[0] += g3[1] * b1[2] - g3[2] * b1[1]; g2[1] += g3[2] * b1[0] - g3[0] * b1[2]; g2[2] += g3[0] * b1[1] - g3[1] * b1[0];
  // b2 = u/|u|
  const float g2b2 = g2[0] * b2[0] + g2[1] * b2[1] + g2[2] * b2[2];
  const float gu[3] = {(g2[0] - g2b2 * b2[0]) * i2, (g2[1] - g2b2 * b2[1]) * i2, (g2[2] - g2b2 * b2[2]) * i2};
  // u = a2 - (b1.a2) b1
  const float gub1 = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ga[3 + k] = gu[k] - gub1 * b1[k];
    g1[k] += -d * gu[k] - gub1 * a[3 + k];
  }
  // b1 = a1/|a1|
  const float g1b1 = g1[0] * b1[0] + g1[1] * b1[1] + g1[2] * b1[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) ga[k] = (g1[k] - g1b1 * b1[k]) * i1;
}

// forward of one chain for all 24 joints from a row in GLOBAL memory (strided per-lane reads; the backward kernel
// is latency-tolerant and keeps its shared memory for the gradient staging)
template <bool KEEP>
__device__ __forceinline__ void chain_forward_row(const float* __restrict__ row, float (&R)[kJ][9], float (&L)[kJ][9],
                                                  float (&P)[kJ][3]) {
  constexpr int parents[kJ] = TCD_PARENTS;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;
  P[0][0] = row[4]; P[0][1] = row[5]; P[0][2] = row[6];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int p = parents[j];
    if (has_child[j]) {
      float a[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) a[k] = row[7 + j * 6 + k];
      rot6d_to_rows_fast(a, L[j]);
      if (p < 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[j][k] = L[j][k];
      } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            R[j][r * 3 + c] = R[p][r * 3] * L[j][c] + R[p][r * 3 + 1] * L[j][3 + c] + R[p][r * 3 + 2] * L[j][6 + c];
      }
    }
    if (p >= 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
        P[j][r] = R[p][r * 3] * c_off[j][0] + R[p][r * 3 + 1] * c_off[j][1] + R[p][r * 3 + 2] * c_off[j][2] + P[p][r];
    }
  }
}

template <bool L1>
__global__ void __launch_bounds__(kLossBwdWarps * 32) loss_backward_kernel(
    const float* __restrict__ model_out, const float* __restrict__ target, const float* __restrict__ p2w, float gscale,
    float* __restrict__ grad, int B, int S, int dn, int tiles_per_sample, int total_tiles) {
  __shared__ float smem[kLossBwdWarps * (32 * kChunkStride + 32 * 12 + 32 * 4)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wt = blockIdx.x * kLossBwdWarps + warp;
  if (wt >= total_tiles) return;
  float* gbuf = smem + warp * (32 * kChunkStride + 32 * 12 + 32 * 4);   // chain gradients of the current chunk
  float* sfeet = gbuf + 32 * kChunkStride;
  float* scont = sfeet + 32 * 12;
  const int rows_per_sample = S * dn;
  const int b = wt / tiles_per_sample, tile = wt - b * tiles_per_sample;
  const int adv = 32 - 2 * dn;
  const int row = tile * adv - dn + lane;                  // row of this lane inside the sample (may be out of range)
  const bool exists = row >= 0 && row < rows_per_sample;
  const bool own = exists && lane >= dn && lane < 32 - dn;
  const float* gm = model_out + (int64_t)b * rows_per_sample * kC;
  const float* gt = target + (int64_t)b * rows_per_sample * kC;
  const float pw = p2w ? p2w[b] : 1.0f;
  const float invB = gscale / (float)B;
  constexpr float two = L1 ? 1.0f : 2.0f;                  // d|d| = sign d, d d^2 = 2 d (loss_del carries the rest)
  const float c_rec = 0.636f * pw * two * invB / ((float)rows_per_sample * kC);
  const float c_vel = 2.964f * pw * two * invB / ((float)(S - 1) * dn * 147.f);
  const float c_fk = 0.646f * pw * two * invB / ((float)rows_per_sample * 69.f);
  const float c_foot = 10.942f * two * invB / ((float)rows_per_sample * 12.f);

  float R[kJ][9], L[kJ][9], P[kJ][3];
  float gP[kJ][3];
#pragma unroll
  for (int j = 0; j < kJ; ++j) gP[j][0] = gP[j][1] = gP[j][2] = 0.f;
  // ---- target chain -> FK gradient seeds (own rows), then the prediction chain (kept)
  if (own) {
    float Pt[kJ][3];
    chain_forward_row<false>(gt + (int64_t)row * kC, R, L, Pt);
    chain_forward_row<true>(gm + (int64_t)row * kC, R, L, P);
#pragma unroll
    for (int j = 1; j < kJ; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float g = c_fk * loss_del<L1>((P[j][c] - P[0][c]) - (Pt[j][c] - Pt[0][c]));
        gP[j][c] = g;
        gP[0][c] -= g;
      }
  } else if (exists) {
    chain_forward_row<true>(gm + (int64_t)row * kC, R, L, P);   // halo: only its feet are needed
  }
  // ---- foot term: exchange feet and contact flags of all 32 rows
  {
    const int fj[4] = {7, 8, 10, 11};
#pragma unroll
    for (int f = 0; f < 4; ++f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) sfeet[lane * 12 + f * 3 + c] = exists ? P[fj[f]][c] : 0.f;
      scont[lane * 4 + f] = exists ? gm[(int64_t)row * kC + f] : 0.f;
    }
    __syncwarp();
    if (own) {
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float g = 0.f;
          const float mine = sfeet[lane * 12 + f * 3 + c];
          if (row + dn < rows_per_sample && scont[lane * 4 + f] > 0.95f)        // pair (row, row+dn): v = next - mine
            g -= c_foot * loss_del<L1>(sfeet[(lane + dn) * 12 + f * 3 + c] - mine);
          if (row - dn >= 0 && scont[(lane - dn) * 4 + f] > 0.95f)              // pair (row-dn, row): v = mine - prev
            g += c_foot * loss_del<L1>(mine - sfeet[(lane - dn) * 12 + f * 3 + c]);
          gP[fj[f]][c] += g;
        }
    }
    __syncwarp();
  }
  // ---- reverse sweep over the tree: gradients w.r.t. world rotations, then local rotations
  float gR[kJ][9];
#pragma unroll
  for (int j = 0; j < kJ; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) gR[j][k] = 0.f;
  constexpr int parents[kJ] = TCD_PARENTS;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;

  auto emit_chunk = [&](int col0, int nc, int j0, int nj) {
    // gbuf[lane][*] holds this lane's chain gradients for columns [col0, col0+nc); add the column-wise terms and store
    __syncwarp();
    if (lane < nc) {
      const int c = col0 + lane;
      for (int r = dn; r < 32 - dn; ++r) {
        const int rr = tile * adv - dn + r;
        if (rr < 0 || rr >= rows_per_sample) continue;
        const float m0 = __ldg(gm + (int64_t)rr * kC + c), t0 = __ldg(gt + (int64_t)rr * kC + c);
        float g = gbuf[r * kChunkStride + lane] + c_rec * loss_del<L1>(m0 - t0);
        if (c >= 4) {
          if (rr + dn < rows_per_sample)
            g -= c_vel * loss_del<L1>((__ldg(gm + (int64_t)(rr + dn) * kC + c) - m0) - (__ldg(gt + (int64_t)(rr + dn) * kC + c) - t0));
          if (rr - dn >= 0)
            g += c_vel * loss_del<L1>((m0 - __ldg(gm + (int64_t)(rr - dn) * kC + c)) - (t0 - __ldg(gt + (int64_t)(rr - dn) * kC + c)));
        }
        grad[((int64_t)b * rows_per_sample + rr) * kC + c] = g;
      }
    }
    __syncwarp();
  };

  // joints in descending order, grouped by the forward chunks (20-23 | 15-19 | 10-14 | 5-9 | 0-4), header last
#pragma unroll
  for (int chunk = 4; chunk >= 0; --chunk) {
    const int j0 = chunk * 5, nj = chunk == 4 ? 4 : 5;
#pragma unroll
    for (int jj = nj - 1; jj >= 0; --jj) {
      const int j = j0 + jj;
      const int p = parents[j];
      float ga[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (own) {
        if (p >= 0) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            gP[p][r] += gP[j][r];
#pragma unroll
            for (int c = 0; c < 3; ++c) gR[p][r * 3 + c] += gP[j][r] * c_off[j][c];
          }
        }
        if (has_child[j]) {
          float gL[9];
          if (p >= 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                // R_j = R_p L_j :  gR_p += gR_j L_j^T ,  gL_j = R_p^T gR_j
                gR[p][r * 3 + c] += gR[j][r * 3] * L[j][c * 3] + gR[j][r * 3 + 1] * L[j][c * 3 + 1] + gR[j][r * 3 + 2] * L[j][c * 3 + 2];
                gL[r * 3 + c] = R[p][r] * gR[j][c] + R[p][3 + r] * gR[j][3 + c] + R[p][6 + r] * gR[j][6 + c];
              }
          } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) gL[k] = gR[j][k];
          }
          float a[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) a[k] = gm[(int64_t)row * kC + 7 + j * 6 + k];
          rot6d_backward(a, gL, ga);
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) gbuf[lane * kChunkStride + jj * 6 + k] = ga[k];
    }
    emit_chunk(7 + j0 * 6, nj * 6, j0, nj);
  }
  // header: contact channels have no chain gradient (the 0.95 threshold is piecewise constant), root = gP[0]
#pragma unroll
  for (int k = 0; k < 4; ++k) gbuf[lane * kChunkStride + k] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) gbuf[lane * kChunkStride + 4 + k] = own ? gP[0][k] : 0.f;
  emit_chunk(0, 7, 0, 0);
}

// second stage: fixed-order sums -> per-sample means * p2w -> batch means * weights.  One block of 32 warps: warp w takes the
// samples w, w + 32, ...; its lanes stride over the sample's tiles (one 16-byte partial each) and meet in a butterfly, so the
// summation order is fixed by the shape alone.  (r01: one warp walked every partial with dependent 4-byte loads, 256 per lane
// at batch 128.)
constexpr int kFinalizeWarps = 32;
__global__ void __launch_bounds__(kFinalizeWarps * 32) loss_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ p2w,
                                                                            float* __restrict__ out, int B, int tiles, int S, int dn) {
  __shared__ float s_acc[4][kFinalizeWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float n_rec = (float)S * dn * kC, n_vel = (float)(S - 1) * dn * 147.f;
  const float n_fk = (float)S * dn * 69.f, n_foot = (float)S * dn * 12.f;
  const float4* part4 = reinterpret_cast<const float4*>(partial);
  for (int b = warp; b < B; b += kFinalizeWarps) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = lane; t < tiles; t += 32) {
      const float4 v = part4[(int64_t)b * tiles + t];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x = warp_sum(s.x); s.y = warp_sum(s.y); s.z = warp_sum(s.z); s.w = warp_sum(s.w);
    const float w = p2w ? p2w[b] : 1.0f;
    acc[0] += s.x / n_rec * w;
    acc[1] += s.y / n_vel * w;
    acc[2] += s.z / n_fk * w;
    acc[3] += s.w / n_foot;
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) s_acc[k][warp] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float wgt[4] = {0.636f, 2.964f, 0.646f, 10.942f};
    float tot = 0.f;
    for (int k = 0; k < 4; ++k) {
      float s = 0.f;
      for (int l = 0; l < kFinalizeWarps; ++l) s += s_acc[k][l];
      float v = wgt[k] * (s / (float)B);
      out[1 + k] = v;
      tot += v;
    }
    out[0] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Post-sampling stage (model/diffusion.py:811-838,942-955 and the long-mode stitching :841-915):
// un-normalise (clip, -min, /scale: dataset/preprocess.py:39-43, dataset/scaler.py:80-83), split contact,
// 6D -> axis-angle (pytorch3d route) and the faithful quaternion FK chain, one thread per (frame, dancer).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float unnorm1(const float* __restrict__ row, const float* __restrict__ mn,
                                         const float* __restrict__ sc, int c) {
  const float v = fminf(fmaxf(__ldg(row + c), -1.0f), 1.0f);
  return __fdiv_rn(__fsub_rn(v, __ldg(mn + c)), __ldg(sc + c));
}
__device__ __forceinline__ void token_joint_aa(const float* row, const float* mn, const float* sc, int j, float* aa) {
  float a[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = unnorm1(row, mn, sc, 7 + 6 * j + k);
  float b1[3], b2[3], b3[3];
  rot6d_to_rows(a, b1, b2, b3);
  quat_to_aa(rows_to_quat(b1, b2, b3), aa);
}
// one step of SMPLSkeleton.forward (vis.py:358-406) for joint j given its local axis-angle
__device__ __forceinline__ void fk_step(int j, const float* aa, const float* root, Quat (&rw)[kJ], float (&pw)[kJ][3]) {
  constexpr int parents[kJ] = TCD_PARENTS;
  constexpr int has_child[kJ] = TCD_HAS_CHILD;
  const Quat ql = aa_to_quat(aa);
  const int p = parents[j];
  if (p < 0) {
    rw[j] = ql;
    pw[j][0] = root[0]; pw[j][1] = root[1]; pw[j][2] = root[2];
  } else {
    Quat pq{0.f, c_off[j][0], c_off[j][1], c_off[j][2]};
    Quat conj{rw[p].w, -rw[p].x, -rw[p].y, -rw[p].z};
    Quat r = qmul_raw(qmul_raw(rw[p], pq), conj);
    pw[j][0] = r.x + pw[p][0]; pw[j][1] = r.y + pw[p][1]; pw[j][2] = r.z + pw[p][2];
    if (has_child[j]) {
      Quat m = qmul_raw(rw[p], ql);
      if (m.w < 0.f) { m.w = -m.w; m.x = -m.x; m.y = -m.y; m.z = -m.z; }
      rw[j] = m;
    }
  }
}

__global__ void __launch_bounds__(128) samples_to_poses_kernel(const float* __restrict__ samples, const float* __restrict__ mn,
                                                               const float* __restrict__ sc, float* __restrict__ contact,
                                                               float* __restrict__ trans, float* __restrict__ aa_out,
                                                               float* __restrict__ joints, int B, int S, int dn) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * S * dn) return;
  const int b = (int)(t / ((int64_t)S * dn)), r = (int)(t % ((int64_t)S * dn)), s = r / dn, d = r % dn;
  const float* row = samples + t * 151;
  const int64_t perm = ((int64_t)b * dn + d) * S + s;                 // (b, dancer, frame)
  if (contact) {
#pragma unroll
    for (int c = 0; c < 4; ++c) contact[perm * 4 + c] = unnorm1(row, mn, sc, c);
  }
  float root[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) root[c] = unnorm1(row, mn, sc, 4 + c);
  if (trans) { trans[t * 3] = root[0]; trans[t * 3 + 1] = root[1]; trans[t * 3 + 2] = root[2]; }
  Quat rw[kJ];
  float pw[kJ][3];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    float aa[3];
    token_joint_aa(row, mn, sc, j, aa);
    if (aa_out) { aa_out[t * 72 + j * 3] = aa[0]; aa_out[t * 72 + j * 3 + 1] = aa[1]; aa_out[t * 72 + j * 3 + 2] = aa[2]; }
    fk_step(j, aa, root, rw, pw);
    if (joints) {
      float* o = joints + (perm * kJ + j) * 3;
      o[0] = pw[j][0]; o[1] = pw[j][1]; o[2] = pw[j][2];
    }
  }
}

// dataset/quaternion.py:35-71 on one quaternion pair
__device__ __forceinline__ Quat slerp1(Quat x, Quat y, float a) {
  float len = x.w * y.w + x.x * y.x + x.y * y.y + x.z * y.z;
  if (len < 0.f) { len = -len; y.w = -y.w; y.x = -y.x; y.y = -y.y; y.z = -y.z; }
  float a0, a1;
  if ((1.0f - len) < 0.01f) {
    a0 = 1.0f - a; a1 = a;
  } else {
    const float om = acosf(len), so = sinf(om);
    a0 = sinf((1.0f - a) * om) / so;
    a1 = sinf(a * om) / so;
  }
  return Quat{a0 * x.w + a1 * y.w, a0 * x.x + a1 * y.x, a0 * x.y + a1 * y.y, a0 * x.z + a1 * y.z};
}

// long mode: the W windows (W, S*dn, 151) of one song overlap by half a window; thread = (output frame f, dancer d)
__global__ void __launch_bounds__(128) samples_to_poses_long_kernel(const float* __restrict__ samples, const float* __restrict__ mn,
                                                                    const float* __restrict__ sc, const float* __restrict__ fade_out,
                                                                    const float* __restrict__ fade_in, const float* __restrict__ sw,
                                                                    float* __restrict__ trans, float* __restrict__ aa_out,
                                                                    float* __restrict__ joints, int W, int S, int dn) {
  const int half = S / 2, F = S + half * (W - 1);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)F * dn) return;
  const int f = (int)(t / dn), d = (int)(t % dn);
  const int blk = f / half, k = f % half;
  const bool overlap = blk >= 1 && blk <= W - 1;
  const int w_left = blk == 0 ? 0 : blk - 1, s_left = blk == 0 ? k : half + k;      // window / frame of the first source
  const float* rl = samples + ((int64_t)w_left * S * dn + (int64_t)s_left * dn + d) * 151;
  const float* rr = samples + ((int64_t)blk * S * dn + (int64_t)k * dn + d) * 151;   // second source (overlap only)
  float root[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = unnorm1(rl, mn, sc, 4 + c);
    if (overlap) v = __fadd_rn(__fmul_rn(v, __ldg(fade_out + k)), __fmul_rn(unnorm1(rr, mn, sc, 4 + c), __ldg(fade_in + k)));
    root[c] = v;
    if (trans) trans[t * 3 + c] = v;
  }
  Quat rw[kJ];
  float pw[kJ][3];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    float aa[3];
    token_joint_aa(rl, mn, sc, j, aa);
    if (overlap) {
      float ar[3];
      token_joint_aa(rr, mn, sc, j, ar);
      quat_to_aa(slerp1(aa_to_quat(aa), aa_to_quat(ar), __ldg(sw + k)), aa);
    }
    if (aa_out) { aa_out[t * 72 + j * 3] = aa[0]; aa_out[t * 72 + j * 3 + 1] = aa[1]; aa_out[t * 72 + j * 3 + 2] = aa[2]; }
    fk_step(j, aa, root, rw, pw);
    if (joints) {
      float* o = joints + (((int64_t)d * F + f) * kJ + j) * 3;
      o[0] = pw[j][0]; o[1] = pw[j][1]; o[2] = pw[j][2];
    }
  }
}

}  // namespace tcd
using namespace tcd;

extern "C" int tcd_samples_to_poses(const float* samples, const float* min_, const float* scale, float* contact, float* trans,
                                    float* poses_aa, float* joints, int B, int S, int dn, void* stream) {
  TCD_REQUIRE(B >= 0 && S > 0 && dn > 0, "tcd_samples_to_poses: bad shape");
  if (B == 0) return TCD_OK;
  TCD_REQUIRE(samples && min_ && scale, "tcd_samples_to_poses: null pointer");
  const int64_t n = (int64_t)B * S * dn;
  samples_to_poses_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(samples, min_, scale, contact, trans, poses_aa, joints,
                                                                          B, S, dn);
  return check_launch("samples_to_poses");
}

extern "C" int tcd_samples_to_poses_long(const float* samples, const float* min_, const float* scale, const float* fade_out,
                                         const float* fade_in, const float* slerp_weight, float* trans, float* poses_aa,
                                         float* joints, int windows, int S, int dn, void* stream) {
  TCD_REQUIRE(windows >= 1 && S > 0 && S % 2 == 0 && dn > 0, "tcd_samples_to_poses_long: bad shape");
  TCD_REQUIRE(samples && min_ && scale && fade_out && fade_in && slerp_weight, "tcd_samples_to_poses_long: null pointer");
  const int64_t n = (int64_t)(S + (S / 2) * (windows - 1)) * dn;
  samples_to_poses_long_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(samples, min_, scale, fade_out, fade_in,
                                                                               slerp_weight, trans, poses_aa, joints, windows, S, dn);
  return check_launch("samples_to_poses_long");
}

extern "C" int tcd_ax_from_6v(const float* d6, float* aa, int64_t n, void* stream) {
  TCD_REQUIRE(n == 0 || (d6 && aa), "tcd_ax_from_6v: null pointer");
  if (n == 0) return TCD_OK;
  ax_from_6v_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d6, aa, n);
  return check_launch("ax_from_6v");
}

extern "C" int tcd_smpl_fk(const float* aa, const float* root, float* pos, int64_t n, void* stream) {
  TCD_REQUIRE(n == 0 || (aa && root && pos), "tcd_smpl_fk: null pointer");
  if (n == 0) return TCD_OK;
  smpl_fk_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(aa, root, pos, n);
  return check_launch("smpl_fk");
}

extern "C" int tcd_motion_fk(const float* motion, float* pos, int64_t n, int C, void* stream) {
  TCD_REQUIRE(C == kC, "tcd_motion_fk: C must be 151, got %d", C);
  TCD_REQUIRE(n == 0 || (motion && pos), "tcd_motion_fk: null pointer");
  if (n == 0) return TCD_OK;
  motion_fk_kernel<<<ceil_div(n, 32), 128, 0, as_stream(stream)>>>(motion, pos, n);
  return check_launch("motion_fk");
}

static int loss_tiles_per_sample(int S, int dn) { return ceil_div((int64_t)S * dn, 32 - dn); }

extern "C" int64_t tcd_loss_workspace_floats(int B, int S, int dn) {
  if (dn < 1 || dn > 16) return 0;
  return (int64_t)B * loss_tiles_per_sample(S, dn) * 4;
}

extern "C" int tcd_loss_backward(const float* model_out, const float* target, const float* p2w, float grad_total,
                                 float* grad_model_out, int B, int S, int dn, int loss_type, void* stream) {
  TCD_REQUIRE(loss_type == TCD_LOSS_L2 || loss_type == TCD_LOSS_L1, "tcd_loss_backward: loss_type must be TCD_LOSS_L2 or TCD_LOSS_L1");
  TCD_REQUIRE(model_out && target && grad_model_out, "tcd_loss_backward: null pointer");
  TCD_REQUIRE(B > 0 && S > 1 && dn > 0 && dn <= 10, "tcd_loss_backward: bad shape B=%d S=%d dn=%d (dancers <= 10)", B, S, dn);
  const int tiles = ceil_div((int64_t)S * dn, 32 - 2 * dn);
  const int64_t total = (int64_t)B * tiles;
  TCD_REQUIRE(total < (1LL << 31), "tcd_loss_backward: too many tiles");
  if (loss_type == TCD_LOSS_L1)
    loss_backward_kernel<true><<<ceil_div(total, kLossBwdWarps), kLossBwdWarps * 32, 0, as_stream(stream)>>>(
        model_out, target, p2w, grad_total, grad_model_out, B, S, dn, tiles, (int)total);
  else
    loss_backward_kernel<false><<<ceil_div(total, kLossBwdWarps), kLossBwdWarps * 32, 0, as_stream(stream)>>>(
        model_out, target, p2w, grad_total, grad_model_out, B, S, dn, tiles, (int)total);
  return check_launch("loss_backward");
}

extern "C" int tcd_loss_forward(const float* model_out, const float* target, const float* p2w, float* workspace,
                                float* losses_out, int B, int S, int dn, int loss_type, void* stream) {
  TCD_REQUIRE(loss_type == TCD_LOSS_L2 || loss_type == TCD_LOSS_L1, "tcd_loss_forward: loss_type must be TCD_LOSS_L2 or TCD_LOSS_L1");
  TCD_REQUIRE(model_out && target && workspace && losses_out, "tcd_loss_forward: null pointer");
  TCD_REQUIRE((uintptr_t)workspace % 16 == 0, "tcd_loss_forward: workspace must be 16-byte aligned");
  TCD_REQUIRE(B > 0 && S > 1 && dn > 0 && dn <= 16, "tcd_loss_forward: bad shape B=%d S=%d dn=%d (dancers <= 16)", B, S, dn);
  const int tiles = loss_tiles_per_sample(S, dn);
  const int64_t total = (int64_t)B * tiles;
  TCD_REQUIRE(total < (1LL << 31), "tcd_loss_forward: too many tiles");
  if (loss_type == TCD_LOSS_L1)
    loss_forward_kernel<true><<<ceil_div(total, kLossWarps), kLossWarps * 32, 0, as_stream(stream)>>>(model_out, target, workspace, S,
                                                                                                      dn, tiles, (int)total);
  else
    loss_forward_kernel<false><<<ceil_div(total, kLossWarps), kLossWarps * 32, 0, as_stream(stream)>>>(model_out, target, workspace, S,
                                                                                                       dn, tiles, (int)total);
  int rc = check_launch("loss_forward");
  if (rc) return rc;
  loss_finalize_kernel<<<1, kFinalizeWarps * 32, 0, as_stream(stream)>>>(workspace, p2w, losses_out, B, tiles, S, dn);
  return check_launch("loss_finalize");
}
