cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc2q -c 2 -o gpurun_out/r02_attn_final -f python tools/kernel_bench.py attn --ncu > gpurun_out/r02_attn_final_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_frn -c 3 -o gpurun_out/r02_frn_final -f python tools/kernel_bench.py fused --ncu > gpurun_out/r02_frn_final_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_attn_final.ncu-rep gpurun_out/r02_frn_final.ncu-rep > gpurun_out/r02_ncu_final_summary.txt 2>&1
tail -5 gpurun_out/r02_ncu_final_summary.txt
