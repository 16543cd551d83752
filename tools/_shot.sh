cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "convert or conditioning or loss" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph 2>&1 | grep -v Warn | tail -1 | cut -c1-260
