#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
# third A/B round: converged issue loops in the training kernels (TCD_TRAIN_CONV)
cd "$(dirname "$0")/.."
export TCD_ATTN_VAR=39 TCD_GEMM_VAR=1 TCD_FRN_VAR=2
for c in 0 1; do
echo "=== TCD_TRAIN_CONV=$c"
if [ $c = 1 ]; then TCD_TRAIN_CONV=$c timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "train_tc or wgrad" 2>&1 | tail -2; fi
TCD_TRAIN_CONV=$c timeout 200 python tools/kernel_bench.py train 2>&1 | tail -12
done
