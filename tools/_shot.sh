cd /root/repo
timeout 100 python tools/kernel_bench.py fused 2>&1 | grep -v Warn | grep tail | cut -c1-250
