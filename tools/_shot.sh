cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -x -q -p no:cacheprovider -k "gemm or attention or fused or tail" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
for l in libtcdiff_sm100a libtcdiff_ab_nopdl; do
echo "== $l"
timeout 300 python tools/kernel_bench.py sampler --lib tcdiff_b200/lib/$l.so 2>&1 | grep -i "sampler c2" | cut -c1-160
timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph --lib tcdiff_b200/lib/$l.so 2>&1 | grep -v Warn | tail -1 | cut -c1-200
done
done
