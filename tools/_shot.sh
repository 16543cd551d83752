cd /root/repo
for l in libtcdiff_sm100a libtcdiff_ab_lb5 libtcdiff_ab_lb6; do
echo "== $l"
timeout 300 python tools/kernel_bench.py loss --lib tcdiff_b200/lib/$l.so 2>&1 | grep loss_forward | cut -c1-100
done
