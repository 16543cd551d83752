cd /root/repo
timeout 300 python -m pytest tests/test_gpu_fused_tail.py -q -p no:cacheprovider 2>&1 | tail -8
