cd /root/repo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_bench_c2_final2.csv python bench.py --steps 1 --warmup 1 --no-train --no-c4 --no-c5 --no-cpu-baseline --no-breakdown > gpurun_out/r02_bench_ncu.log 2>&1
python tools/launch_agg.py gpurun_out/r02_launches_bench_c2_final2.csv --top 12
