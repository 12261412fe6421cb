set -x
nvidia-smi -L
python -m pytest tests -m gpu -q -x --timeout 1200 2>&1 | tail -25
for k in 4 3; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel $k > gpurun_out/bench_k$k.json 2> gpurun_out/bench_k$k.err; tail -c 1500 gpurun_out/bench_k$k.json; done
for w in 9 7 5; do python bench.py --steps 4 --warmup 3 --no-cpu-baseline --kernel 4 --tile-warps $w > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; python -c "
import json;d=json.load(open('gpurun_out/bench_w$w.json'));print('W=$w',d['value'],d['e2e']['value'])"; done
for c in C1 C4 C5 C3; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$c.json'));print('$c',d['value'],d['e2e']['value'],d['roofline']['frac'])"; done
