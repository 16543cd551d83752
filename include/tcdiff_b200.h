/* tcdiff_b200.h — C-ABI of libtcdiff_sm100a.so: hand-written sm_100a (B200) kernels for the TCDiff
 * denoising hot path.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every entry returns 0 (TCD_OK) or a negative code; tcd_last_error() gives the message
 *     (thread-local);
 *   - all pointers are DEVICE pointers unless named host_*; tensors are contiguous row-major
 *     unless a leading dimension / stride argument is given (counted in ELEMENTS);
 *   - `stream` is a cudaStream_t passed as void*; no entry allocates or synchronises, so every
 *     entry is CUDA-graph capturable;
 *   - `dtype` selects the operand element type of GEMM/attention inputs: TCD_F32 (parity mode,
 *     CUDA-core fp32 FMA) or TCD_BF16 (tcgen05 tensor cores, fp32 accumulate).  The residual
 *     stream, LayerNorm statistics, softmax, schedule math and the DDIM/DDPM update are fp32
 *     in both modes.
 *
 * Reference file:line citations are relative to the TCDiff repository root.
 */
#ifndef TCDIFF_B200_H_
#define TCDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCD_OK 0
#define TCD_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define TCD_ERR_CUDA (-2)    /* CUDA runtime / driver error */

enum { TCD_F32 = 0, TCD_BF16 = 1 };
enum { TCD_ACT_NONE = 0, TCD_ACT_RELU = 1, TCD_ACT_GELU = 2, TCD_ACT_MISH = 3, TCD_ACT_SILU = 4,
       TCD_ACT_LEAKY_RELU = 5 /* nn.LeakyReLU(0.01): TrajDecoder MLPs */ };
enum { TCD_LOSS_L2 = 0 /* F.mse_loss */, TCD_LOSS_L1 = 1 /* F.l1_loss (model/diffusion.py:172) */ };

const char* tcd_last_error(void);
/* ABI version and compiled architecture ("sm_100a"). */
int tcd_version(void);
const char* tcd_arch(void);
/* Compile-time tuning choices of this build (csrc/tuning.cuh): "fuse_tails", "attn_2q", "frn_rc", "fold_ln"; -1 for an unknown name.  There is
 * no run-time switch of any kind: the host reads what the library was built with. */
int tcd_tuning(const char* name);

/* ------------------------------------------------------------------------------------------------
 * Diffusion step kernels (fp32).  x, outputs, noise: (n_tokens, C) with C = 151; traj: (n_tokens, 3).
 * ---------------------------------------------------------------------------------------------- */

/* Classifier-free-guidance blend + x0 clamp + eps-from-x0 + stochastic DDIM update + trajectory
 * in-painting, one pass.  Replaces DanceDecoder.guided_forward's blend (model/model.py:546),
 * GaussianDiffusion.model_predictions / predict_noise_from_start (model/diffusion.py:189-204) and the
 * body of ddim_sample's loop (model/diffusion.py:411-431):
 *   o = unc + (con - unc) * w;  x0 = clip ? clamp(o,-1,1) : o;  eps = (sqrt_recip*x - x0)/sqrt_recipm1
 *   last ? x' = x0 : x' = x0*sqrt_alpha_next + c*eps + sigma*noise;   x'[..., 4:6] = traj[..., 0:2]
 * traj may be NULL (no in-painting, x_0=None).  x0_out (optional) receives x0.  xpad_out (optional)
 * receives x' as bf16 rows of pitch `xpad_ld` (>= C, zero padded) — the A operand of the next step's
 * input projection.  noise may be NULL when last != 0.  x_out may alias x. */
int tcd_cfg_ddim_step(const float* x, const float* out_cond, const float* out_uncond, const float* noise,
                      const float* traj, float* x_out, float* x0_out, void* xpad_out, int64_t xpad_ld,
                      int64_t n_tokens, int C, float w, float sqrt_recip, float sqrt_recipm1,
                      float sqrt_alpha_next, float c, float sigma, int clip, int last, void* stream);

/* CFG blend + clamp + posterior mean + noise for ancestral sampling.  Replaces p_mean_variance /
 * q_posterior / p_sample (model/diffusion.py:206-252):
 *   o = unc + (con-unc)*w;  predict_epsilon != 0: o = sqrt_recip*x - sqrt_recipm1*o (predict_start_from_noise, :176-187);
 *   x0 = clamp(o);  x' = coef1*x0 + coef2*x + nonzero * std * noise,
 * std = exp(0.5*posterior_log_variance_clipped[t]).  mask/value (optional, both or neither) implement
 * inpaint_loop's constraint (model/diffusion.py:547-549): x' = value_q*mask + (1-mask)*x'. */
int tcd_cfg_ddpm_step(const float* x, const float* out_cond, const float* out_uncond, const float* noise,
                      float* x_out, void* xpad_out, int64_t xpad_ld, int64_t n_tokens, int C, float w,
                      float coef1, float coef2, float std, int nonzero, int predict_epsilon, float sqrt_recip,
                      float sqrt_recipm1, const float* mask, const float* value_q, void* stream);

/* The same two step kernels with the step's Gaussian draw generated in the kernel instead of read from a noise tensor
 * (model/diffusion.py:421 `noise = torch.randn_like(x)`, :246 in p_sample): element i of draw `rng_stream` is component
 * i & 3 of Philox4x32-10 block i >> 2 (Box-Muller), keyed by the device-resident pair rng_state = uint64 {seed, call
 * counter}.  tcd_philox_normal writes exactly that draw to memory (x_T, q_sample draws of inpaint_loop, and the parity
 * tests, which feed it to the noise-tensor entries above and require bit-identical results). */
int tcd_cfg_ddim_step_rng(const float* x, const float* out_cond, const float* out_uncond, const void* rng_state,
                          uint32_t rng_stream, const float* traj, float* x_out, float* x0_out, void* xpad_out,
                          int64_t xpad_ld, int64_t n_tokens, int C, float w, float sqrt_recip, float sqrt_recipm1,
                          float sqrt_alpha_next, float c, float sigma, int clip, int last, void* stream);
int tcd_cfg_ddpm_step_rng(const float* x, const float* out_cond, const float* out_uncond, const void* rng_state,
                          uint32_t rng_stream, float* x_out, void* xpad_out, int64_t xpad_ld, int64_t n_tokens, int C,
                          float w, float coef1, float coef2, float std, int nonzero, int predict_epsilon,
                          float sqrt_recip, float sqrt_recipm1, const float* mask, const float* value_q, void* stream);
int tcd_philox_normal(float* out, int64_t n, const void* rng_state, uint32_t rng_stream, void* stream);

/* x[..., 4:6] = traj[..., 0:2] (model/diffusion.py:396-403,434-440); optional bf16 padded copy. */
int tcd_inpaint_traj(float* x, const float* traj, void* xpad_out, int64_t xpad_ld, int64_t n_tokens, int C,
                     void* stream);

/* Training-time forward diffusion of p_losses (model/diffusion.py:640-651): x_start is (B, dn, S, C);
 * writes x_noisy as (B, S, dn, C) = sqrt_ac[t_b]*x + sqrt_1mac[t_b]*noise with channels 4,5 restored
 * from x_start, and target (B,S,dn,C) = permuted x_start (optional).  noise is (B,S,dn,C).  t: int64 (B).
 * Also covers q_sample (model/diffusion.py:625-634) with permute=0 and restore_traj=0. */
int tcd_q_sample(const float* x_start, const float* noise, const int64_t* t, const float* sqrt_ac,
                 const float* sqrt_1mac, float* x_noisy, float* target, void* xpad_out, int64_t xpad_ld,
                 int B, int dn, int S, int C, int permute, int restore_traj, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kinematics and loss (fp32).
 * ---------------------------------------------------------------------------------------------- */

/* 6D rotation -> axis-angle.  Replaces ax_from_6v (dataset/quaternion.py:28-32 -> pytorch3d 0.7.1
 * rotation_6d_to_matrix + matrix_to_axis_angle).  d6: (n,6), aa: (n,3). */
int tcd_ax_from_6v(const float* d6, float* aa, int64_t n, void* stream);

/* SMPL forward kinematics over the 24-joint tree.  Replaces SMPLSkeleton.forward (vis.py:358-406).
 * aa: (n,24,3) axis-angle local rotations, root: (n,3), pos: (n,24,3) world joint positions. */
int tcd_smpl_fk(const float* aa, const float* root, float* pos, int64_t n, void* stream);

/* Fused 6D -> joints: pos (n,24,3) from motion rows (n, C=151) laid out [contact4, root3, rot6d 24x6]
 * (dataset/group_dataset.py:213).  Equivalent to ax_from_6v + SMPLSkeleton.forward on rows
 * (model/diffusion.py:692-708) via a direct rotation-matrix chain. */
int tcd_motion_fk(const float* motion, float* pos, int64_t n, int C, void* stream);

/* Workspace floats needed by tcd_loss_forward (the per-block partial sums of the four loss terms,
 * model/diffusion.py:664-741). */
int64_t tcd_loss_workspace_floats(int B, int S, int dn);
/* The four p_losses terms (model/diffusion.py:664-741): model_out, target (B,S,dn,151) — the target is x_start, or the
 * noise when the reference is built with predict_epsilon=True (:657-660) —, p2w (B) = p2_loss_weight[t], loss_type =
 * TCD_LOSS_L2 (F.mse_loss; TCDiff.py:90-102) or TCD_LOSS_L1 (F.l1_loss, the reference constructor's default, :172).
 * losses_out[0..4] = {total, 0.636*recon, 2.964*vel, 0.646*fk, 10.942*foot}.  Deterministic two-stage reduction
 * through `workspace` (16-byte aligned; one {recon, vel, fk, foot} partial per 32-row tile). */
int tcd_loss_forward(const float* model_out, const float* target, const float* p2w, float* workspace,
                     float* losses_out, int B, int S, int dn, int loss_type, void* stream);

/* Gradient of the same objective w.r.t. the network output: grad_model_out (B,S,dn,151) =
 * grad_total * d total / d model_out (reverse-mode through the 24-joint chain and the 6-D Gram-Schmidt map; the
 * contact > 0.95 mask and the target carry no gradient).  What autograd computes for model/diffusion.py:664-741. */
int tcd_loss_backward(const float* model_out, const float* target, const float* p2w, float grad_total,
                      float* grad_model_out, int B, int S, int dn, int loss_type, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Denoiser building blocks.
 * ---------------------------------------------------------------------------------------------- */

/* C = act(A * W^T + bias).  A: (M,K) pitch lda, W: (N,K) pitch ldw (nn.Linear weight layout), both of
 * `dtype`; bias fp32 (N) or NULL; C: (M,N) pitch ldc of `out_dtype` (TCD_F32 or `dtype`).
 * TCD_F32: CUDA-core fp32 FMA GEMM.  TCD_BF16: tcgen05.mma (TMEM accumulators) fed by TMA; requires
 * 16-byte aligned A/W base pointers and pitches (lda, ldw multiples of 8 elements).
 * Replaces every nn.Linear on the path (model/model.py:60-64,197-199,272-274,294,454-465,474,
 * 490-501,519,522-528). */
int tcd_gemm(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int act,
             int out_dtype, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream);

/* LayerNorm (+ optional rotary) producing GEMM operands.  x: (rows, D) fp32.  out_plain / out_rot
 * (either may be NULL) of `dtype`: LN(x) and rotary(LN(x)) with position = row % tokens_per_sample,
 * cos/sin tables (>= tokens_per_sample, D/2) fp32.  Replaces norm1-4 + rotate_queries_or_keys
 * (model/model.py:326,331-332,338,344,375,387; model/rotary_embedding_torch.py:39-59,107-130). */
int tcd_layernorm_rotary(int dtype, const float* x, const float* gamma, const float* beta, float eps,
                         void* out_plain, void* out_rot, const float* rot_cos, const float* rot_sin,
                         int64_t rows, int D, int tokens_per_sample, void* stream);

/* Standalone rotary embedding of fp32 rows (RotaryEmbedding.rotate_queries_or_keys,
 * model/rotary_embedding_torch.py:107-113): out = x*cos + rotate_half(x)*sin, position = row % tokens_per_sample. */
int tcd_rotary(const float* x, float* out, const float* rot_cos, const float* rot_sin, int64_t rows, int D,
               int tokens_per_sample, void* stream);

/* Fused block tail: x_out = x_in + (1 + scale) * LNin(y) + shift, then operands of the next block.
 *   x_in/x_out: (rows, D) fp32 residual stream (x_out may equal x_in; x_out may be NULL when only the next
 *   LayerNorm of the updated x is consumed: the feed-forward tail, whose x is replaced by linear3(norm4(x)),
 *   model/model.py:344,371);
 *   y: (rows, D) of y_dtype (GEMM output);  ln_in_* optional inner LayerNorm (SBI_MSA.layer_norm,
 *   eps 1e-6, model/model.py:68,106);  film: (samples, film_ld) fp32 rows holding [scale(D) | shift(D)]
 *   at column offset film_off, or NULL for a plain residual add (music encoder, model/model.py:219-220);
 *   next_* optional LayerNorm (+rotary) of the updated x -> out_plain / out_rot of `dtype`.
 * Replaces featurewise_affine + residual (model/model.py:171-173,327,334,339). */
int tcd_film_residual_norm(int dtype, const float* x_in, float* x_out, const void* y, int y_dtype,
                           const float* ln_in_gamma, const float* ln_in_beta, float ln_in_eps, const float* film,
                           int64_t film_ld, int64_t film_off, const float* next_gamma, const float* next_beta,
                           float next_eps, void* out_plain, void* out_rot, const float* rot_cos,
                           const float* rot_sin, int64_t rows, int D, int tokens_per_sample, void* stream);

/* tcd_gemm (bf16, N = D = 512) + tcd_film_residual_norm in ONE kernel: y = A W^T (+ bias) never leaves the SM
 * (fp32 accumulator in tensor memory), v = x_in + (1 + scale) * LNin(y) + shift, x_out = v, operands of the next block
 * = LNnext(v) (+ rotary) in bf16.  A (M,K) bf16 pitch lda, W (512,K) bf16 pitch ldw (nn.Linear weight as stored), bias (512)
 * fp32 or NULL; x_in (M,512) fp32 contiguous or NULL (no residual), x_out likewise (may equal x_in) or NULL (not written);
 * ln_in_gamma / beta NULL = no inner LayerNorm; film NULL = no modulation (v = x_in + LNin(y)); next_gamma / next_beta
 * both NULL = LNnext without its affine (out_plain only, attention / feed-forward tail shapes: the caller has folded gamma /
 * beta into the weights of the nn.Linear that reads out_plain, model/model.py:338-339,344); at least one of out_plain /
 * out_rot (M,512) bf16; for out_rot the rotary tables TRANSPOSED, rot_cos_t / rot_sin_t
 * (256 angles, rot_ld >= tokens_per_sample positions) fp32, so that the 32 rows of a warp read them coalesced.  A cluster of two CTAs owns a 128-row tile and splits it by
 * columns (two tensor-memory accumulators per CTA: the tail of tile i overlaps the MMAs of tile i+1), row statistics are
 * exchanged through distributed shared memory (csrc/gemm_frn.cu).  Same reference lines as tcd_film_residual_norm plus the
 * `fc` / `linear2` projections (model/model.py:64,103,274,400).  Which decoder tails the sampler runs through this entry is
 * the compile-time choice TCD_TUNE_FUSE_TAILS (csrc/tuning.cuh). */
int tcd_gemm_film_residual_norm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                int64_t M, int64_t K, const float* x_in, float* x_out,
                                const float* ln_in_gamma, const float* ln_in_beta, float ln_in_eps,
                                const float* film, int64_t film_ld, int64_t film_off,
                                const float* next_gamma, const float* next_beta, float next_eps,
                                void* out_plain, void* out_rot, const float* rot_cos_t, const float* rot_sin_t,
                                int64_t rot_ld, int tokens_per_sample, void* stream);

/* Profiling aid of tcd_gemm_film_residual_norm: buf = 8 device uint64 cycle counters (summed over CTAs and launches; epilogue
 * warp: [0] waiting for the MMAs, [1] pass 1, [2] pass 2, [3] second exchange, [4] pass 3, [5] waiting for residual boxes,
 * [6] waiting in the statistics exchanges; MMA warp: [7] waiting for the epilogue) or NULL to switch the instrumentation off.
 * No reference counterpart (the reference runs model/model.py:327,334,339 as separate aten ops). */
int tcd_gemm_frn_set_debug(void* buf);

/* softmax(scale * Q K^T) V per (sample, head), head_dim 64, no mask.  Q: rows of pitch ldq holding
 * heads at column h*64; same for K, V, O.  Replaces SBI_MSA's core (model/model.py:97-102) and the
 * nn.MultiheadAttention core of the music encoder (model/model.py:232-239). */
int tcd_attention(int dtype, const void* Q, int64_t ldq, int64_t q_batch_stride, const void* K, int64_t ldk,
                  int64_t k_batch_stride, const void* V, int64_t ldv, int64_t v_batch_stride, void* O,
                  int64_t ldo, int64_t o_batch_stride, int samples, int heads, int Lq, int Lk, float scale,
                  void* stream);

/* SinusoidalPosEmb (model/utils.py:41-48) as a gather from a host-built (n_timestep, D) fp32 table (built
 * with the reference's exact fp32 expression): times int64 (n) -> (n, D) of `dtype`. */
int tcd_time_embed(int dtype, const int64_t* times, const float* table, void* out, int n, int D, int n_timestep,
                   void* stream);

/* Music-token post-processing (model/model.py:585-597): tokens (n,S,D) fp32 are replaced in place by
 * null_embed (S,D) where keep[b]==0; pooled_ln (n,D) of `dtype` = LayerNorm(mean over S) with
 * gamma/beta (non_attn_cond_projection.0). */
int tcd_cond_pool(int dtype, float* tokens, const float* null_embed, const uint8_t* keep, const float* gamma,
                  const float* beta, void* pooled_ln, int n, int S, int D, void* stream);

/* t = t_lin + (keep ? cond_hidden : null_hidden) (model/model.py:609-612); writes t (n,D) fp32 and
 * Mish(t) of `dtype` (the DenseFiLM input, model/model.py:161,165). */
int tcd_time_cond(int dtype, const float* t_lin, const float* cond_hidden, const float* null_hidden,
                  const uint8_t* keep, float* t_out, void* mish_out, int n, int D, void* stream);

/* Hoisted sampler form of the same: every step s shares one timestep across the batch, so
 * mish_out[s, j, :] = Mish(t_lin[s] + (j < B ? ch_cond[j] : ch_uncond)) for the 2B conditional+unconditional
 * rows of step s (model/model.py:542-546 runs both passes with the same times). */
int tcd_sampler_time_cond(int dtype, const float* t_lin, const float* ch_cond, const float* ch_uncond,
                          void* mish_out, int steps, int B, int D, void* stream);

/* Cross-attention memory (model/model.py:615-616 + rotary of :388): mem = norm_cond(cat(tokens (n,S,D),
 * t_tokens (n,2,D))) -> mem_plain, mem_rot (n,S+2,D) of `dtype` (rotary position = memory index). */
int tcd_build_memory(int dtype, const float* tokens, const float* t_tokens, const float* gamma,
                     const float* beta, void* mem_plain, void* mem_rot, const float* rot_cos,
                     const float* rot_sin, int n, int S, int D, void* stream);

/* Strided row copy: dst[b, dst_row0 + r, :cols] = src[b * src_batch_stride + r * src_ld + :cols] for r < rows,
 * b < samples (src_batch_stride = 0 broadcasts one source block to every sample).  Used by the hoisted sampler
 * to refresh the two time-token rows of the cross-attention K/V buffers, to duplicate the shared front for the
 * unconditional pass, and for the window hand-over of long_ddim_sample / long_inpaint_loop
 * (x[1:, :half] = x[:-1, half:], model/diffusion.py:502-506,599-601).  src and dst must not overlap. */
int tcd_scatter_rows(int dtype, const void* src, int64_t src_ld, int64_t src_batch_stride, void* dst, int64_t dst_ld,
                     int64_t dst_batch_stride, int64_t dst_row0, int rows, int cols, int samples, void* stream);

/* x[b,s,d,c] = w[s,c]*value[b,s,d,c] + (1-w[s,c])*x[b,s,d,c] with a (S, C) fp32 weight table shared by batch and
 * dancer: the per-step hard overwrite (w in {0,1}) and the final linear-interpolation fusion of
 * ddim_sample_Footwork (model/diffusion.py:307-309,343-344,371-379). */
int tcd_masked_blend(float* x, const float* value, const float* weight, int B, int S, int dn, int C, void* stream);

/* fp32 -> dtype conversion with zero padding: dst (rows, dst_ld) <- src (rows, src_ld)[:, :cols].  Packs nn.Linear
 * weights and the (.., 151) motion rows fed to input_projection (model/model.py:474,561; `cond_embed.float()` :578) as
 * K-padded GEMM operands. */
int tcd_convert_pad(int dtype, const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                    int cols, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward-pass primitives (fp32 gradients).  Together with tcd_gemm (dgrad: dY W, wgrad: dY^T X through
 * tcd_cast_transpose), tcd_rotary (rotation by -theta) and tcd_attention_backward they are what autograd executes
 * for the reference's training step (model/diffusion.py:636-753, TCDiff.py:227-234).
 * ---------------------------------------------------------------------------------------------- */

/* Refreshes every bf16 operand copy of the training step in ONE launch.  `segments` is a device array of n_segments records
 * { const float* src; bf16* w; bf16* wt; int64 src_ld, w_ld, wt_ld; int32 rows, cols; } (56 bytes): w[r, c] = bf16(src[r, c]),
 * wt[c, r] = the same value (wt may be NULL).  Pad rows / columns of the destinations are not written.  Replaces the
 * per-parameter `.to(bf16)` autocast copies a mixed-precision run of TCDiff.py:227-245 would make, and this repo's own
 * per-weight tcd_convert_pad + tcd_cast_transpose launches (234 per step in round 1). */
int tcd_pack_weights(const void* segments, int n_segments, int blocks_per_segment, void* stream);

/* dst (cols, rows) of `dtype` = transpose(src (rows, cols) fp32); pitches in elements (operand of the nn.Linear weight
 * gradient dW = dY^T X that autograd forms for every nn.Linear of model/model.py:60-64,197-199,272-294,456-527). */
int tcd_cast_transpose(int dtype, const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows, int64_t cols,
                       void* stream);
/* out[g, c] (+)= sum over the rows_per_group rows of group g of a[row, c] * (b ? b[row, c] : 1)  (bias gradients,
 * LayerNorm dgamma/dbeta partials, FiLM): the reductions autograd performs for nn.Linear.bias, nn.LayerNorm.weight/bias
 * (model/model.py:68,202-203,277-279,296,471) and DenseFiLM (model/model.py:164-168). */
int tcd_group_colsum(const float* a, const float* b, int64_t ld, int64_t groups, int64_t rows_per_group, int cols,
                     float* out, int64_t out_ld, int accumulate, void* stream);
/* y = act(z) and dx = dy * act'(z) for TCD_ACT_{RELU,GELU,MISH,SILU}: F.relu / nn.ReLU (model/model.py:492,524,526),
 * F.gelu(erf) (:427 -> :244,400), nn.Mish (:161,457), nn.SiLU (:499). */
int tcd_act_forward(int act, const float* z, float* y, int64_t n, void* stream);
int tcd_act_backward(int act, const float* z, const float* dy, float* dx, int64_t n, void* stream);
/* nn.LayerNorm backward (model/model.py:68,202-203,277-279,296,471,497) over rows of D: dx, and per-warp partial sums dgamma_part/dbeta_part of shape
 * (tcd_layernorm_backward_partials(rows), D) to be reduced with tcd_group_colsum. */
int64_t tcd_layernorm_backward_partials(int64_t rows);
int tcd_layernorm_backward(const float* x, const float* gamma, const float* dy, float eps, float* dx, float* dgamma_part,
                           float* dbeta_part, int64_t rows, int D, void* stream);
/* featurewise_affine + residual backward (model/model.py:171-173): out = x + (1+scale[b]) v + shift[b]:
 * dv = (1+scale) dout, dfilm[b, off:off+D] = sum_rows dout*v, dfilm[b, off+D:off+2D] = sum_rows dout (dx = dout). */
int tcd_film_backward(const float* dout, const float* v, const float* film, int64_t film_ld, int64_t film_off, float* dv,
                      float* dfilm, int64_t dfilm_ld, int64_t dfilm_off, int samples, int L, int D, void* stream);

/* Attention backward (fp32) of SBI_MSA's core (model/model.py:97-102) and of the music encoder's nn.MultiheadAttention
 * (:232-239): dQ, dK, dV of O = softmax(scale Q K^T) V per (sample, head), head_dim 64; flash-style
 * (P recomputed from a log-sum-exp pass; workspace = tcd_attention_backward_workspace_floats floats).  Row layouts as
 * in tcd_attention (pitch, batch stride; heads at column h*64). */
int64_t tcd_attention_backward_workspace_floats(int samples, int heads, int Lq);
int tcd_attention_backward(const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs,
                           const float* V, int64_t ldv, int64_t vbs, const float* O, int64_t ldo, int64_t obs,
                           const float* dO, int64_t ldg, int64_t gbs, float* dQ, int64_t lddq, int64_t dqbs, float* dK,
                           int64_t lddk, int64_t dkbs, float* dV, int64_t lddv, int64_t dvbs, float* workspace,
                           int samples, int heads, int Lq, int Lk, float scale, void* stream);

/* bf16 training tape (GEMM operands and their gradients in bf16, residual stream and parameter gradients fp32):
 * the mixed-precision counterparts of the fp32 backward primitives above. */
/* Each of the following kernels can apply one nn.Dropout site for free (dropout_p > 0; mask of tcd_dropout indexed by
 * the flat element index of the named operand): act -> the activation's OUTPUT (model.py:400 FFN inner dropout);
 * layernorm_bf16 / its backward -> the norm's INPUT (model.py:103); film -> v (model.py:383,396,401 and 240,245). */
int tcd_act_forward_bf16(int act, const void* z, void* y, int64_t n, float dropout_p, const void* rng_state, uint32_t site,
                         void* stream);
int tcd_act_backward_bf16(int act, const void* z, const void* dy, void* dx, int64_t n, float dropout_p,
                          const void* rng_state, uint32_t site, void* stream);
/* LayerNorm backward with upstream gradients dy (and optionally dy_rot, the gradient of the rotary copy produced by
 * tcd_layernorm_rotary, rotated back by -theta and added) of dy_dtype; x, dx and the optional dres (gradient arriving
 * through the residual connection around the norm, added to dx) of x_dtype; dgamma_part / dbeta_part receive
 * tcd_layernorm_backward_mixed_partials(rows) fp32 partial rows each (one per thread block), to be summed with
 * tcd_group_colsum.  dy may be NULL when only dy_rot flows.  (x, dy) dtypes: (f32, bf16), (bf16, bf16), (f32, f32). */
int64_t tcd_layernorm_backward_mixed_partials(int64_t rows);
int tcd_layernorm_backward_mixed(int x_dtype, int dy_dtype, const void* x, const float* gamma, const void* dy,
                                 const void* dy_rot, const float* rot_cos, const float* rot_sin, int tokens_per_sample,
                                 float eps, const void* dres, void* dx, float* dgamma_part, float* dbeta_part, int64_t rows,
                                 int D, float x_dropout_p, const void* rng_state, uint32_t site, void* stream);
/* nn.LayerNorm over rows of D with bf16 input and output (SBI_MSA.layer_norm on the bf16 fc output,
 * model/model.py:68,103-105). */
int tcd_layernorm_bf16(const void* x, const float* gamma, const float* beta, float eps, void* y, int64_t rows, int D,
                       float x_dropout_p, const void* rng_state, uint32_t site, void* stream);
/* tcd_film_backward (featurewise_affine + residual, model/model.py:171-173,327,334,339) with bf16 v / dv (dout, film,
 * dfilm fp32); workspace: tcd_film_backward_workspace_floats. */
int64_t tcd_film_backward_workspace_floats(int samples, int L, int D);
int tcd_film_backward_bf16(const float* dout, const void* v, const float* film, int64_t film_ld, int64_t film_off, void* dv,
                           float* dfilm, int64_t dfilm_ld, int64_t dfilm_off, float* workspace, int samples, int L, int D,
                           float v_dropout_p, const void* rng_state, uint32_t site, void* stream);
/* forward of the same block on the training tape (model/model.py:327,334,339; film NULL = the music encoder's plain
 * residual :219-220): out = x + (1 + scale[b]) v + shift[b] with bf16 v (film NULL: x + v). */
int tcd_film_residual_bf16(const float* x, const void* v, const float* film, int64_t film_ld, int64_t film_off, float* out,
                           int64_t rows, int L, int D, float v_dropout_p, const void* rng_state, uint32_t site, void* stream);
/* out[c] = sum over rows of a[row, c] for a bf16 (rows, cols) matrix with pitch ld (gradients of the nn.Linear biases,
 * e.g. model/model.py:197-199,272-274,294,474,519).  ld even and `a` 4-byte aligned; a pitch that is a multiple of 8
 * elements on a 16-byte aligned base takes the 16-byte-load kernel. */
int64_t tcd_colsum_bf16_workspace_floats(int64_t rows, int cols);
int tcd_colsum_bf16(const void* a, int64_t ld, int64_t rows, int cols, float* out, float* workspace, void* stream);

/* nn.Dropout for the training tape (sites model/model.py:98,103,240,244-245,383,396,400-401): y = x * keep / (1-p) with a counter-based mask (csrc/dropout.cuh) derived from the
 * device-resident rng_state = {uint64 seed, uint64 step counter}, the call-site id and the flat element index; the same
 * call on dy is the backward pass.  Nothing is stored; a CUDA-graph replay sees the updated counter. */
int tcd_dropout(int dtype, const void* x, void* y, int64_t n, float p, const void* rng_state, uint32_t site, void* stream);
/* The attention-probability mask the training attention kernels apply, materialised as fp32 (samples, heads, Lq, Lk)
 * with values 0 or 1/(1-p) (tests inject it into the oracle). */
int tcd_dropout_mask_attention(float* out, int samples, int heads, int Lq, int Lk, float p, const void* rng_state,
                               uint32_t site, void* stream);

/* Weight-gradient contraction on the tcgen05 tensor cores (what `accelerator.backward(total_loss)`, TCDiff.py:232,
 * computes for every nn.Linear.weight): C (M,N) fp32 = A^T B with A (K,M) and B (K,N) bf16
 * row-major as stored (rows = tokens), i.e. dW = dY^T X of nn.Linear without transposing the activations; split
 * along K over the SMs with a deterministic second-pass reduction.  workspace: tcd_gemm_tn_workspace_floats floats,
 * 16-byte aligned; lda, ldb multiples of 8. */
int64_t tcd_gemm_tn_workspace_floats(int64_t M, int64_t N, int64_t K);
int tcd_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K, float* workspace, void* stream);

/* bf16 attention for the training step on the tcgen05 tensor cores (SBI_MSA core model/model.py:97-102 and its
 * autograd).  Forward = tcd_attention(TCD_BF16) that also
 * writes lse[sample, head, q] = log2-domain log-sum-exp of the scaled scores; backward recomputes P from it
 * (flash-style, deterministic, no atomics): dQ, dK, dV in bf16.  All matrices bf16 with the row layouts of
 * tcd_attention; stats_ws holds tcd_attention_train_workspace_floats floats (8-byte aligned).
 * dropout_p > 0: dropout on the attention probabilities (model/model.py:98; nn.MultiheadAttention dropout) with the
 * counter-based mask of tcd_dropout_mask_attention — forward and backward must get the same (rng_state, site). */
int64_t tcd_attention_train_workspace_floats(int samples, int heads, int Lq);
int tcd_attention_train_forward(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs,
                                const void* V, int64_t ldv, int64_t vbs, void* O, int64_t ldo, int64_t obs, float* lse,
                                int samples, int heads, int Lq, int Lk, float scale, float dropout_p, const void* rng_state,
                                uint32_t site, void* stream);
int tcd_attention_train_backward(const void* Q, int64_t ldq, int64_t qbs, const void* K, int64_t ldk, int64_t kbs,
                                 const void* V, int64_t ldv, int64_t vbs, const void* O, int64_t ldo, int64_t obs,
                                 const void* dO, int64_t ldg, int64_t gbs, const float* lse, void* dQ, int64_t lddq,
                                 int64_t dqbs, void* dK, int64_t lddk, int64_t dkbs, void* dV, int64_t lddv, int64_t dvbs,
                                 float* stats_ws, int samples, int heads, int Lq, int Lk, float scale, float dropout_p,
                                 const void* rng_state, uint32_t site, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Post-sampling stage ("next" row N2): what render_sample does between the sampler and the renderer
 * (model/diffusion.py:811-838,942-955; long mode :841-915).
 * ---------------------------------------------------------------------------------------------- */

/* samples (B, S*dn, 151) normalised, frame-major tokens -> un-normalise with the MinMaxScaler's min_/scale_ (151 each;
 * clip to [-1,1], subtract, divide: dataset/preprocess.py:39-43, dataset/scaler.py:80-83), then
 *   contact (B, dn, S, 4) | trans (B, S*dn, 3) | poses_aa (B, S*dn, 24, 3) axis-angle via the pytorch3d route |
 *   joints (B, dn, S, 24, 3) = SMPLSkeleton.forward (vis.py:358-406).  Any output may be NULL. */
int tcd_samples_to_poses(const float* samples, const float* min_, const float* scale, float* contact, float* trans,
                         float* poses_aa, float* joints, int B, int S, int dn, void* stream);
/* long mode: the `windows` half-overlapping windows of ONE song stitched (root positions cross-faded with
 * fade_out/fade_in (S/2 floats each = linspace(1,0) / linspace(0,1)), joint rotations slerped with slerp_weight
 * (dataset/quaternion.py:35-71)), F = S + S/2 (windows-1) frames: trans (F, dn, 3) | poses_aa (F, dn, 24, 3) |
 * joints (dn, F, 24, 3). */
int tcd_samples_to_poses_long(const float* samples, const float* min_, const float* scale, const float* fade_out,
                              const float* fade_in, const float* slerp_weight, float* trans, float* poses_aa,
                              float* joints, int windows, int S, int dn, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Trajectory front end of the end-to-end test mode ("next" row N3: TrajDecoder/model/traj_model.py:125-200,
 * TrajDecoder/utils/utils_model.py:10-74, driver loop TCDiff.py:526-556).  A small fp32 model (hidden 64 / 128):
 * its GEMMs / LayerNorms / residuals reuse tcd_gemm(TCD_F32) / tcd_layernorm_rotary / tcd_film_residual_norm.
 * ---------------------------------------------------------------------------------------------- */

/* One nn.LSTM layer (gate order i, f, g, o; zero initial state), hidden size 64, input size I <= 64:
 * x[t, n, :I] at x + t*x_ts + n*x_ld, out[t, n, :64] at out + t*out_ts + n*out_ld (+ add_table[n, :64] if not NULL:
 * the batch-first PositionalEncoding the reference adds to the last layer's output, traj_model.py:101).
 * The recurrence runs over t = 0..T-1 (the reference feeds (b, dn*seq, c) to a batch_first=False LSTM, so t is the
 * BATCH axis, traj_model.py:139,174) — one thread block per n. */
int tcd_lstm_layer(const float* x, int64_t x_ts, int64_t x_ld, const float* w_ih, const float* w_hh, const float* b_ih,
                   const float* b_hh, float* out, int64_t out_ts, int64_t out_ld, const float* add_table, int T, int N,
                   int I, void* stream);
/* tcd_attention(TCD_F32) with head_dim 32 or 64 (TrajDecoder: 4 heads of 32; CausalCrossConditionalSelfAttention.forward,
 * TrajDecoder/model/traj_model.py:29-47 — despite its name it applies no mask). */
int tcd_attention_f32_hd(int head_dim, const float* Q, int64_t ldq, int64_t qbs, const float* K, int64_t ldk, int64_t kbs,
                         const float* V, int64_t ldv, int64_t vbs, float* O, int64_t ldo, int64_t obs, int samples, int heads,
                         int Lq, int Lk, float scale, void* stream);
/* kalman_smooth_batch (utils_model.py:10-74; filterpy 1.4.5 KalmanFilter, constant-velocity model): per track
 * x <- F x; x <- x + K_t (z_t - H x) in float64 with the data-independent gain sequence K_t (T x 4 x 2 doubles,
 * computed on the host from F, H, Q, R, P0); xy / out (tracks, T, 2) fp32. */
int tcd_kalman_smooth(const float* xy, float* out, const double* gains, int tracks, int T, double dt, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step of the data-parallel training loop over flat fp32 arenas ("next" row N1).
 * ---------------------------------------------------------------------------------------------- */

/* One Adan update (model/adan.py:33-123, restart_cond=None) of `count` parameters plus, when ema != NULL, the
 * EMA blend of the master model with the UPDATED parameters (model/diffusion.py:61-76, TCDiff.py:242-245):
 * ema = ema*ema_beta + (1-ema_beta)*param.  `step` is the number of updates already applied (state["step"]): the
 * first call (step 0) leaves the moments untouched like the reference (adan.py:70) and only applies the weight-decay
 * division.  grad is multiplied by grad_scale first (1/world after a SUM all-reduce; 1 = untouched).  All arenas are
 * 16-byte aligned. */
int tcd_adan_ema_step(float* param, const float* grad, float* prev_grad, float* exp_avg, float* exp_avg_diff,
                      float* exp_avg_sq, float* ema, int64_t count, int64_t step, double grad_scale, double lr,
                      double beta1, double beta2, double beta3, double eps, double weight_decay, double ema_beta,
                      void* stream);
/* Same update (model/adan.py:33-123 + model/diffusion.py:61-76) with the step counter resident on the device (*step_device is read, used as `step`, and incremented by a
 * one-thread prologue kernel that also derives the bias corrections), so the call can be captured in a CUDA graph and
 * replayed.  scalars_workspace: 128 bytes of device memory, 16-byte aligned. */
int tcd_adan_ema_step_device(float* param, const float* grad, float* prev_grad, float* exp_avg, float* exp_avg_diff,
                             float* exp_avg_sq, float* ema, int64_t count, int64_t* step_device, void* scalars_workspace,
                             double grad_scale, double lr, double beta1, double beta2, double beta3, double eps,
                             double weight_decay, double ema_beta, void* stream);
/* ema = ema*beta + (1-beta)*param (model/diffusion.py:73-76) for parameters the optimizer does not touch. */
int tcd_ema_update(float* ema, const float* param, int64_t count, double beta, void* stream);
/* The same blend over a list of tensors in one launch (EMA.update_model_average walks 446 parameter tensors,
 * model/diffusion.py:66-71): ema_ptrs / param_ptrs / counts are DEVICE arrays of n_tensors entries. */
int tcd_ema_update_multi(const void* ema_ptrs, const void* param_ptrs, const int64_t* counts, int n_tensors,
                         int64_t max_count, double beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCDIFF_B200_H_ */
