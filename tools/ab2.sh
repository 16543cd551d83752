#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
# second A/B round: converged issue loops (attention VAR bits 5/6, TCD_GEMM_VAR=1)
cd "$(dirname "$0")/.."
for v in 39 99 103; do
echo "=== TCD_ATTN_VAR=$v"
TCD_ATTN_VAR=$v timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attention" 2>&1 | tail -2
TCD_ATTN_VAR=$v timeout 200 python tools/kernel_bench.py attn 2>&1 | tail -2
done
for g in 0 1; do
echo "=== TCD_GEMM_VAR=$g"
if [ $g = 1 ]; then TCD_GEMM_VAR=$g timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm" 2>&1 | tail -2; fi
TCD_GEMM_VAR=$g timeout 200 python tools/kernel_bench.py gemm 2>&1 | tail -7
done
