#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in sm100a ab_attn_old; do timeout 200 python tools/attn_err.py --lib tcdiff_b200/lib/libtcdiff_$v.so 2>&1 | tail -7; done
TCDIFF_TEST_LIB=tcdiff_b200/lib/libtcdiff_ab_attn_old.so timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q -s -p no:cacheprovider -k "training_gradients_vs_oracle_autograd or training_gradients_with_dropout" 2>&1 | grep -E "cos|passed|failed" | head
