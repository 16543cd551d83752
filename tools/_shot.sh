cd /root/repo
# launch list of the c2 leg of the bench command (graph replay: every kernel node appears as its own launch)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r02_launches_bench_c2.csv python bench.py --steps 1 --warmup 1 --no-train --no-c4 --no-c5 --no-cpu-baseline --no-breakdown > gpurun_out/r02_launches_bench.log 2>&1
tail -2 gpurun_out/r02_launches_bench.log | cut -c1-300
wc -l gpurun_out/r02_launches_bench_c2.csv
