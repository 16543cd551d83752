cd /root/repo
timeout 300 python -m pytest tests/test_gpu_fused_tail.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 60 python tools/kernel_bench.py fused --phases 2>&1 | grep -v Warn | cut -c1-250
