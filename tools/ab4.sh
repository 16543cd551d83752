#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
# fourth A/B round: packed-f32x2 GELU epilogue (1-CTA vs CTA-pair kernel), attention VAR 43
cd "$(dirname "$0")/.."
echo "=== gelu2 default (1-CTA lane-0)"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm" 2>&1 | tail -2
timeout 200 python tools/kernel_bench.py gemm 2>&1 | grep act2
echo "=== gelu2 1-CTA converged"; TCD_GEMM_VAR=1 timeout 200 python tools/kernel_bench.py gemm 2>&1 | grep act2
echo "=== gelu2 pair"; TCD_GEMM_GELU_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm" 2>&1 | tail -2
TCD_GEMM_GELU_PAIR=1 timeout 200 python tools/kernel_bench.py gemm 2>&1 | grep act2
echo "=== attn 43"; TCD_ATTN_VAR=43 timeout 200 python tools/kernel_bench.py attn 2>&1 | tail -2
echo "=== attn 39"; TCD_ATTN_VAR=39 timeout 200 python tools/kernel_bench.py attn 2>&1 | tail -2
