cd /root/repo
for i in 1 2 3; do
for d in build/ab_base .; do
echo "== $d"
(cd $d && timeout 300 python tools/kernel_bench.py sampler 2>&1 | grep -i "sampler c2" | cut -c1-110)
(cd $d && timeout 300 python tools/train_bench.py --steps 10 --warmup 3 --graph 2>&1 | grep -v Warn | tail -1 | cut -c1-190)
done
done
