cd /root/repo
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 400 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
print('train',{k:d['train'].get(k) for k in ('value','ms_per_step','n_gpus')})
print('c4',d['c4']['value'],'c5',d['c5']['value'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
