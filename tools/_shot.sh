cd /root/repo
timeout 900 python -m pytest tests/test_gpu_fused_tail.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -p no:cacheprovider -k "bf16 or c2 or bench_batch or golden" 2>&1 | tail -3
for i in 1 2; do
for l in libtcdiff_sm100a libtcdiff_ab_nofold; do
echo "== $l"
timeout 300 python tools/kernel_bench.py sampler --lib tcdiff_b200/lib/$l.so 2>&1 | grep -i "sampler\|clips" | cut -c1-200
done
done
timeout 300 python tools/kernel_bench.py fused 2>&1 | tail -12 | cut -c1-220
