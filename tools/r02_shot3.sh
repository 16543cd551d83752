#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB=tcdiff_b200/lib/libtcdiff_ab_attn2q.so
TCDIFF_TEST_LIB=$AB timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 300 -k "attention" > gpurun_out/s3_attn_tests.log 2>&1; tail -15 gpurun_out/s3_attn_tests.log
for v in sm100a ab_attn2q; do
  timeout 400 python tools/kernel_bench.py attn sampler --lib tcdiff_b200/lib/libtcdiff_$v.so > gpurun_out/s3_kb_$v.log 2>&1
  echo "== $v"; grep -E "library|attn|sampler" gpurun_out/s3_kb_$v.log | cut -c1-250
done

