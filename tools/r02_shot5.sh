#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q -s -p no:cacheprovider -k "training_gradients_vs_oracle_autograd or training_gradients_with_dropout or network_backward_bf16" > gpurun_out/s5_tests.log 2>&1; grep -E "^E |whole|worst|passed|failed|Error" gpurun_out/s5_tests.log | head -30
