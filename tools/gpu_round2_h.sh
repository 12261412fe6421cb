set -x
python tools/gpu_parity_diag.py > gpurun_out/parity_diag_v4.txt 2>&1
grep "kernel 4" gpurun_out/parity_diag_v4.txt | cut -c1-200
for c in C2 C1 C4 C5; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $c --kernel 4 > gpurun_out/bench_v4_$c.json 2> gpurun_out/bench_v4_$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v4_$c.json'));print('$c',d['value'],d['e2e']['value'],d['roofline']['frac'])"; done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config C3 --kernel 4 > gpurun_out/bench_v4_C3.json 2> gpurun_out/bench_v4_C3.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v4_C3.json'));print('C3',d['value'],d['e2e']['value'],d['roofline']['frac'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err && python -c "
import json;d=json.load(open('gpurun_out/plain_small.json'));print('small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_v4_C2_b4736 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log | cut -c1-100
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -15
