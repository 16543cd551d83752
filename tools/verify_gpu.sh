#!/bin/bash
# full GPU verification on one B200: test suite, smoke(), kernel micro-benchmarks of the hot kernels
cd "$(dirname "$0")/.."
(time timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -4) 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 200 python tools/kernel_bench.py attn gemm frn 2>&1 | tail -12
