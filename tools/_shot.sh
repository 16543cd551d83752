cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "attention" 2>&1 | tail -2
for l in libtcdiff_ab_base libtcdiff_sm100a libtcdiff_sm100a libtcdiff_ab_base libtcdiff_ab_base libtcdiff_sm100a; do
echo "== $l"
timeout 300 python tools/kernel_bench.py attn --lib tcdiff_b200/lib/$l.so 2>&1 | grep -i "attn n128" | cut -c1-100
timeout 300 python tools/kernel_bench.py sampler --lib tcdiff_b200/lib/$l.so 2>&1 | grep -i "sampler c2" | cut -c1-110
done
