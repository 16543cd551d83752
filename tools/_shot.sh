cd /root/repo
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 300 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
print('train',{k:d['train'].get(k) for k in ('value','ms_per_step','n_gpus')})
print('c4',d['c4']['value'],'c5',d['c5']['value'])
PY
