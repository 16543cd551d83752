cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider 2>&1 | tail -3
