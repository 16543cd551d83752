#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): the experiments round 1 left built but unmeasured.
#   1. full GPU suite + smoke on the default path (regression gate for everything below)
#   2. GELU epilogue A/B: TCD_GELU_VAR=0|1 on the GELU GEMM shape (kernel_bench "act2" line), then the model / sampler parity
#      tests and the bench line with TCD_GELU_VAR=1.  Adopt only if parity stays green AND the act2 line drops below ~100 us.
#   3. row-chunked tail pairs again with chunk sizes that DIVIDE the 128 samples of a c2 step (round 1 measured 25 and 12, whose
#      last chunks are 3 and 8 samples: part of the loss was the ragged tail, not the idea): TCD_TAIL_CHUNK 64 / 32 / 16.
# Outputs under gpurun_out/r02a_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r02a_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 > gpurun_out/r02a_smoke.log
for v in 0 1 0 1; do
  echo "=== TCD_GELU_VAR=$v"; TCD_GELU_VAR=$v timeout 200 python tools/kernel_bench.py gemm 2>&1 | grep -i "act2\|Error" 
done > gpurun_out/r02a_gelu_kernel.log 2>&1
TCD_GELU_VAR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02a_gelu_tests.log
timeout 400 python bench.py --no-train --no-cpu-baseline > gpurun_out/r02a_bench_gelu0.json 2> gpurun_out/r02a_bench_gelu0.err
TCD_GELU_VAR=1 timeout 400 python bench.py --no-train --no-cpu-baseline > gpurun_out/r02a_bench_gelu1.json 2> gpurun_out/r02a_bench_gelu1.err
timeout 300 python tools/tail_chunk_ab.py 0 64 32 16 > gpurun_out/r02a_tail_chunk.log 2>&1
tail -n 5 gpurun_out/r02a_tail_chunk.log gpurun_out/r02a_tests.log gpurun_out/r02a_smoke.log gpurun_out/r02a_gelu_kernel.log gpurun_out/r02a_gelu_tests.log
cut -c1-160 gpurun_out/r02a_bench_gelu0.json gpurun_out/r02a_bench_gelu1.json
