cd /root/repo
export PYTHONUNBUFFERED=1
(time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 30 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 1400 -x) > gpurun_out/sanitizer_full.log 2>&1; echo "rc=$?"
grep -v "^$" gpurun_out/sanitizer_full.log | tail -25 | cut -c1-300
