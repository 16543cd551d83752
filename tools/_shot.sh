cd /root/repo
for i in 1 2; do
timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph 2>&1 | grep -v Warn | tail -1 | cut -c1-200
timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph --no-direct-grads 2>&1 | grep -v Warn | tail -1 | cut -c1-200
done
