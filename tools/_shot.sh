cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_frn -c 2 -o gpurun_out/r02_frn4 -f python tools/kernel_bench.py fused --ncu > gpurun_out/r02_frn4_ncu.log 2>&1
