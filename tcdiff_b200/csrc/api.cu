// C-ABI glue: error reporting and dtype dispatch for the entries declared in include/tcdiff_b200.h.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "tuning.cuh"

namespace tcd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TCD_ERR_CUDA;
  }
  return TCD_OK;
}

int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, int act, float* C,
             int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st);
int gemm_bf16_tc(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act, int out_dtype,
                 void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, cudaStream_t st);
int attention_f32(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs, const float* V,
                  int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                  float scale, cudaStream_t st);
int attention_bf16_tc(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs, const void* V,
                      int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, int samples, int heads, int Lq, int Lk,
                      float scale, float* lse, float dropout_p, const void* rng_state, uint32_t site, cudaStream_t st);

}  // namespace tcd

using namespace tcd;

extern "C" const char* tcd_last_error(void) { return g_err; }
extern "C" int tcd_version(void) { return 1; }
extern "C" const char* tcd_arch(void) { return "sm_100a"; }

extern "C" int tcd_gemm(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act,
                        int out_dtype, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
  TCD_REQUIRE(M >= 0 && N >= 0 && K > 0, "tcd_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  if (M == 0 || N == 0) return TCD_OK;
  TCD_REQUIRE(A && W && C, "tcd_gemm: null pointer");
  TCD_REQUIRE(lda >= K && ldw >= K && ldc >= N, "tcd_gemm: pitch smaller than extent");
  TCD_REQUIRE(act >= TCD_ACT_NONE && act <= TCD_ACT_LEAKY_RELU, "tcd_gemm: bad activation %d", act);
  if (dtype == TCD_F32) {
    TCD_REQUIRE(out_dtype == TCD_F32, "tcd_gemm(f32): output must be f32");
    return gemm_f32((const float*)A, lda, (const float*)W, ldw, bias, act, (float*)C, ldc, M, N, K, as_stream(stream));
  }
  if (dtype == TCD_BF16) {
    TCD_REQUIRE(out_dtype == TCD_F32 || out_dtype == TCD_BF16, "tcd_gemm(bf16): bad out dtype %d", out_dtype);
    return gemm_bf16_tc(A, lda, W, ldw, bias, act, out_dtype, C, ldc, M, N, K, as_stream(stream));
  }
  set_error("tcd_gemm: bad dtype %d", dtype);
  return TCD_ERR_INVALID;
}

extern "C" int tcd_tuning(const char* name) {
  if (name == nullptr) return -1;
  if (strcmp(name, "fuse_tails") == 0) return TCD_TUNE_FUSE_TAILS;
  if (strcmp(name, "attn_2q") == 0) return TCD_TUNE_ATTN_2Q;
  if (strcmp(name, "frn_rc") == 0) return TCD_TUNE_FRN_RC;
  if (strcmp(name, "fold_ln") == 0) return TCD_TUNE_FOLD_LN;
  if (strcmp(name, "pdl") == 0) return TCD_TUNE_PDL;
  return -1;
}

extern "C" int tcd_attention(int dtype, const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                             int64_t k_batch_stride, const void* V, int64_t ldv, int64_t v_batch_stride, void* O,
                             int64_t ldo, int64_t o_batch_stride, int samples, int heads, int Lq, int Lk, float scale,
                             void* stream) {
  TCD_REQUIRE(samples >= 0 && heads > 0 && Lq >= 0 && Lk > 0, "tcd_attention: bad shape");
  if (samples == 0 || Lq == 0) return TCD_OK;
  TCD_REQUIRE(Q && K && V && O, "tcd_attention: null pointer");
  TCD_REQUIRE(heads <= 65535 && samples <= 65535, "tcd_attention: grid limit");
  if (dtype == TCD_F32) {
    TCD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 &&
                ((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) % 16 == 0 &&
                q_batch_stride % 4 == 0 && k_batch_stride % 4 == 0 && v_batch_stride % 4 == 0 && o_batch_stride % 4 == 0,
                "tcd_attention(f32): 16-byte alignment required");
    return attention_f32((const float*)Q, ldq, q_batch_stride, (const float*)K, ldk, k_batch_stride, (const float*)V,
                         ldv, v_batch_stride, (float*)O, ldo, o_batch_stride, samples, heads, Lq, Lk, scale,
                         as_stream(stream));
  }
  if (dtype == TCD_BF16) {
    return attention_bf16_tc(Q, ldq, q_batch_stride, K, ldk, k_batch_stride, V, ldv, v_batch_stride, O, ldo,
                             o_batch_stride, samples, heads, Lq, Lk, scale, nullptr, 0.f, nullptr, 0u, as_stream(stream));
  }
  set_error("tcd_attention: bad dtype %d", dtype);
  return TCD_ERR_INVALID;
}
