#!/bin/bash
# A/B helper for one B200: kernel_bench (args after --) for the product library and every tcdiff_b200/lib/libtcdiff_ab_*.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in tcdiff_b200/lib/libtcdiff_sm100a.so tcdiff_b200/lib/libtcdiff_ab_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  timeout 600 python tools/kernel_bench.py "$@" --lib "$lib" 2>&1 | grep -vE "Warning|warn" | cut -c1-170
done
