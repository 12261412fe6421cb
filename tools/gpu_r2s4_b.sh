# session 4, call b: one --set full capture of the Dual<4> tile kernel (C3, one tile per SM)
python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 263 > gpurun_out/c3_small.json 2> gpurun_out/c3_small.err && python -c "
import json;d=json.loads(open('gpurun_out/c3_small.json').read().strip().splitlines()[-1]);print('small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_dual4_C3_b263 -f python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 263 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log | cut -c1-200
