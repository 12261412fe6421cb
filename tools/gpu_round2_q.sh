python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --kernel 4 --batch 4736 > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err && python -c "
import json;d=json.load(open('gpurun_out/plain_small.json'));print('small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_v6_C2_b4736 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --kernel 4 --batch 4736 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log | cut -c1-100
