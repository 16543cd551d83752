#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q -s -p no:cacheprovider -k "bench_batch_rows or graphed_train_step or teacher_forced_all_50 or c2_headline" > gpurun_out/s2_tests.log 2>&1; tail -60 gpurun_out/s2_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:film_residual_norm -c 1 -o gpurun_out/s2_frn python tools/kernel_bench.py frn --ncu > gpurun_out/s2_ncu.log 2>&1; tail -3 gpurun_out/s2_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/s2_attn python tools/kernel_bench.py attn --ncu > gpurun_out/s2_ncu2.log 2>&1; tail -3 gpurun_out/s2_ncu2.log
ls -la gpurun_out/*.ncu-rep
