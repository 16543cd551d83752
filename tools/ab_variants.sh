#!/bin/bash
# NOTE: the numbered attention / frn variants these scripts select exist up to commit c621b2d; the current tree keeps
# TCD_ATTN_VAR=0|1, TCD_FRN_VAR=0|1, TCD_GEMM_VAR=0|1|2, TCD_GEMM_GELU_PAIR, TCD_TRAIN_CONV (README.md).
# A/B of the kernel tuning variants (TCD_ATTN_VAR, TCD_FRN_VAR) on one B200: parity tests + micro-benchmarks.
# Usage (GPU box): bash tools/ab_variants.sh > gpurun_out/ab.log 2>&1
cd "$(dirname "$0")/.."
run() {  # attn_var frn_var tests(0/1) benches
  echo "=== TCD_ATTN_VAR=$1 TCD_FRN_VAR=$2"
  if [ "$3" = 1 ]; then
    TCD_ATTN_VAR=$1 TCD_FRN_VAR=$2 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "attention or film_residual" 2>&1 | tail -3
  fi
  TCD_ATTN_VAR=$1 TCD_FRN_VAR=$2 timeout 200 python tools/kernel_bench.py $4 2>&1 | tail -12
}
run 0 0 0 "attn frn"
run 3 1 1 "attn frn"
run 7 2 1 "attn frn"
run 11 1 1 "attn"
run 19 1 1 "attn"
run 23 1 1 "attn"
