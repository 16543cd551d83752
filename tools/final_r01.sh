#!/bin/bash
# round-1 closing run on one B200: full GPU suite, bench line, ncu launch list of the bench command, ncu --set full of the
# attention and GEMM kernels.  Outputs under gpurun_out/ (s2_*).
cd "$(dirname "$0")/.."
(time timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/s2_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 400 --csv --log-file gpurun_out/s2_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/s2_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -c 2 -f -o gpurun_out/s2_attn \
  python tools/kernel_bench.py attn --ncu > gpurun_out/s2_ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -c 1 -f -o gpurun_out/s2_gemm \
  python tools/kernel_bench.py gemm --ncu > gpurun_out/s2_ncu_gemm.log 2>&1
cat gpurun_out/s2_tests.log; cut -c1-300 gpurun_out/s2_bench.json; ls -la gpurun_out/s2_*
