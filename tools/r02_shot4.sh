#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/s4_attn2q python tools/kernel_bench.py attn --ncu --lib tcdiff_b200/lib/libtcdiff_ab_attn2q.so > gpurun_out/s4_ncu.log 2>&1; tail -3 gpurun_out/s4_ncu.log
