cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "attention" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -p no:cacheprovider -k "dropout or gradients or replays" 2>&1 | tail -2
for l in libtcdiff_ab_base libtcdiff_sm100a libtcdiff_sm100a libtcdiff_ab_base libtcdiff_ab_base libtcdiff_sm100a; do
echo "== $l"
timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph --lib tcdiff_b200/lib/$l.so 2>&1 | grep -v Warn | tail -1 | cut -c1-190
done
