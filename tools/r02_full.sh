#!/bin/bash
# full GPU verification on one B200: complete test suite, smoke(), bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -15) > gpurun_out/full_tests.log 2>&1
tail -18 gpurun_out/full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/full_smoke.log 2>&1; tail -3 gpurun_out/full_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -c 600 gpurun_out/full_bench.json
