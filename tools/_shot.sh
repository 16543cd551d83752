cd /root/repo
timeout 300 python -m pytest tests/test_gpu_fused_tail.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 60 python tools/kernel_bench.py fused 2>&1 | grep -v Warn | grep tail | cut -c1-250
