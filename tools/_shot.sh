cd /root/repo
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -3)
for l in libtcdiff_ab_nopdl libtcdiff_sm100a libtcdiff_sm100a libtcdiff_ab_nopdl; do
echo "== $l"
timeout 600 python tools/c4_bench.py --steps 200 --lib tcdiff_b200/lib/$l.so 2>&1 | grep -v Warn | tail -1 | cut -c80-230
timeout 300 python tools/kernel_bench.py sampler --lib tcdiff_b200/lib/$l.so 2>&1 | grep -i "sampler c2" | cut -c1-110
done
