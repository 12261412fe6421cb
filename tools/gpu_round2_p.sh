(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err); tail -5 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu.json').read().strip().split('\n')[-1])
print('C2', d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for e in d.get('extra_configs',[]):
    print(e['config']['workload'][:40], e.get('value'), e['e2e'])
PY
