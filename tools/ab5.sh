#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
cd "$(dirname "$0")/.."
T=tests/test_gpu_train.py::test_training_gradients_with_dropout_vs_oracle
echo "=== default"; timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |assert|passed|failed" | head -20
echo "=== ATTN 35"; TCD_ATTN_VAR=35 timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |passed|failed" | head -8
echo "=== ATTN 0"; TCD_ATTN_VAR=0 timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |passed|failed" | head -8
echo "=== TRAIN_CONV 0"; TCD_TRAIN_CONV=0 timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |passed|failed" | head -8
echo "=== GEMM 0"; TCD_GEMM_VAR=0 TCD_GEMM_GELU_PAIR=0 timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |passed|failed" | head -8
echo "=== FRN 0"; TCD_FRN_VAR=0 timeout 300 python -m pytest $T -x -q 2>&1 | grep -E "^E |passed|failed" | head -8
