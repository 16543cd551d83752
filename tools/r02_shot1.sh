#!/bin/bash
# round-2 first GPU shot: full GPU suite, bench line, A/B builds (GELU rational erf, ring tail) in separate processes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --timeout 600 2>&1 | tail -40) > gpurun_out/s1_tests.log 2>&1
tail -30 gpurun_out/s1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s1_smoke.log 2>&1; tail -4 gpurun_out/s1_smoke.log
for v in sm100a ab_gelu ab_ring; do
  timeout 400 python tools/kernel_bench.py gemm frn sampler --lib tcdiff_b200/lib/libtcdiff_$v.so > gpurun_out/s1_kb_$v.log 2>&1
  echo "== $v"; grep -E "library|N1024 K512 outbf16 act2|film_residual|layernorm_rotary|sampler" gpurun_out/s1_kb_$v.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 3000 gpurun_out/s1_bench.json; tail -5 gpurun_out/s1_bench.err
