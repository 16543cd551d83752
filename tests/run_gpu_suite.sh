#!/bin/bash
# gpurun helper: run the GPU tests in separate processes (a trapped kernel leaves a sticky CUDA error),
# each under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q -p no:cacheprovider --timeout 300 > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; echo "== $name"; tail -n 25 gpurun_out/$name.log; }
run k_misc tests/test_gpu_kernels.py -m gpu -k "not bf16_tcgen05 and not padded_views and not attention"
run k_attn tests/test_gpu_kernels.py -m gpu -k "attention"
run k_gemm tests/test_gpu_kernels.py -m gpu -k "bf16_tcgen05 or padded_views"
run m_fp32 tests/test_gpu_model.py -m gpu -k "fp32 or dead_code or graph_chunks"
run m_bf16 tests/test_gpu_model.py -m gpu -k "bf16"
run train tests/test_gpu_train.py tests/test_gpu_traj.py -m gpu
