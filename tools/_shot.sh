cd /root/repo
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "gemm" 2>&1 | tail -4
bash tools/gpu_ab.sh gemm 2>&1 | grep -E "==|gemm M"
