set -x
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -30
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel 4 > gpurun_out/bench_v2_k4.json 2> gpurun_out/bench_v2_k4.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v2_k4.json'));print('tile C2',d['value'],d['e2e']['value'])"
for w in 10 7; do python bench.py --steps 4 --warmup 3 --no-cpu-baseline --kernel 4 --tile-warps $w > gpurun_out/bench_v2_w$w.json 2> gpurun_out/bench_v2_w$w.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v2_w$w.json'));print('W=$w',d['value'],d['e2e']['value'])"; done
for c in C1 C4 C5; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/bench_v2_$c.json 2> gpurun_out/bench_v2_$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v2_$c.json'));print('$c',d['value'],d['e2e']['value'],d['roofline']['frac'])"; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err && python -c "
import json;d=json.load(open('gpurun_out/plain_small.json'));print('small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_v2_C2_b4736 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
