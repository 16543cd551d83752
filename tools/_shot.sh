cd /root/repo
timeout 120 python tools/kernel_bench.py step 2>&1 | grep -v Warn | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "ddim or step or philox or noise" 2>&1 | tail -3
