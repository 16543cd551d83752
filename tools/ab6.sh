#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
# attention VAR 163 = 35 + deferred epilogue (O double-buffered in TMEM)
cd "$(dirname "$0")/.."
echo "=== tests TCD_ATTN_VAR=163"
TCD_ATTN_VAR=163 timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train.py -q -x -m gpu -k "attention or ddim or forward or gradients or train_step" 2>&1 | tail -3
for v in 163 35 163 35; do echo "=== TCD_ATTN_VAR=$v"; TCD_ATTN_VAR=$v timeout 200 python tools/kernel_bench.py attn 2>&1 | tail -2; done
