cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "loss or fk" 2>&1 | tail -2
for l in libtcdiff_sm100a libtcdiff_ab_pf64 libtcdiff_ab_pf0 libtcdiff_ab_pf128_r32b3; do
echo "== $l"
timeout 300 python tools/kernel_bench.py loss --lib tcdiff_b200/lib/$l.so 2>&1 | grep loss_forward | cut -c1-90
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:loss_forward -c 1 -o gpurun_out/r02_loss4 -f python tools/kernel_bench.py loss --ncu > gpurun_out/r02_loss_ncu.log 2>&1
ncu -i gpurun_out/r02_loss4.ncu-rep --page raw --csv > gpurun_out/r02_loss4_raw.csv 2>/dev/null
