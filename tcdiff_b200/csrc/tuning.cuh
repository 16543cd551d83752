// Compile-time tuning choices of the sm_100a kernels.  Every alternative that was measured on a B200 and lost is gone from
// the tree (history: profiles/r01_*.md, profiles/r02_*.md); what remains here are the few choices still worth an A/B build:
//     python -m tcdiff_b200.build --define TCD_TUNE_ATTN_2Q=0 --out libtcdiff_ab.so
// The product library is built with the defaults below; nothing is switchable at run time (no environment variables).
#pragma once

// Fused fc / linear2 GEMM + FiLM + residual + LayerNorm tails in the sampler (csrc/gemm_frn.cu; engine.py reads tcd_tuning()):
// bit 0 self-attention tail, bit 1 cross-attention tail, bit 2 feed-forward tail.  r02 on a B200, 96 000 rows, fused kernel
// against the unfused GEMM + tail pair: cross-attention 128-132 us / 155 us, self-attention 150 / 161 (rotary factors from
// transposed tables with coalesced loads; 159 with the row-major tables staged through shared memory), feed-forward 132 / 146
// (with a third operand stage in place of one residual slot and the second output buffer; 150 with two stages: K = 1024
// starves them); c2 sampler, same box: 124.5 (0) -> 126.5 (2); 123.2 (2) -> 123.5 (3); 123.3 (3) -> 124.5 (7).
#ifndef TCD_TUNE_FUSE_TAILS
#define TCD_TUNE_FUSE_TAILS 7
#endif

// Attention forward (attention_tc.cu): 1 = one CTA per SM with two 128-query tiles, one softmax thread per query row, P as a
// tensor-memory operand (r02: 295 -> 242 us for the sampler's self-attention launch, 98 -> 88 us cross-attention),
// 0 = two CTAs per SM with one tile each and two threads per row (still used for sequences of at most 128 queries)
#ifndef TCD_TUNE_ATTN_2Q
#define TCD_TUNE_ATTN_2Q 1
#endif

// FiLM + residual + LayerNorm tail (norm.cu): 1 = contiguous row chunks per warp, parameters in registers, input rows through a
// bulk-copy shared-memory ring; 0 = grid-stride rows, one next row prefetched in registers, parameters re-read per row
#ifndef TCD_TUNE_FRN_RC
#define TCD_TUNE_FRN_RC 1
#endif

// LayerNorm affine of norm3 / norm4 folded into the weights of the projection that reads the normalised rows (linear1, linear3:
// W diag(gamma), b + W beta, built once in engine.PackedWeights) — the cross-attention and feed-forward tails of the fused kernel
// then skip LN_next's two per-column vectors (gemm_frn.cu, F_NOAFF).  Only where that tail runs fused (TCD_TUNE_FUSE_TAILS
// bits 1 / 2) and only for plain outputs: the affine does not commute with the rotation of the rotary operands.
#ifndef TCD_TUNE_FOLD_LN
#define TCD_TUNE_FOLD_LN 1
#endif

// Programmatic dependent launch between the persistent tensor-core kernels (GEMM pairs, fused GEMM + tail, two-tile attention)
// for launches of at most ~two tiles per CTA — the launch-bound shapes of BASELINE config 4 (r02: +5.5 %); the long launches of
// the batch-64 sampler and the training step run at the power cap and are launched plainly.  common.cuh, pdl_sync() / launch_pdl().
#ifndef TCD_TUNE_PDL
#define TCD_TUNE_PDL 1
#endif
