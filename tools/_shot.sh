cd /root/repo
timeout 600 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider -k "reference_defaults" 2>&1 | grep -E "^E|assert|passed|failed" | head -12
