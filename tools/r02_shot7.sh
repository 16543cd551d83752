#!/bin/bash
cd "$(dirname "$0")/.."
for v in sm100a; do timeout 300 python tools/grad_cos.py --lib tcdiff_b200/lib/libtcdiff_$v.so 2>&1 | grep -E "library|seed|contact|out rel|Error"; done
