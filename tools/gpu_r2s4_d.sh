# session 4, call d: --set full captures of the tile kernel on C4 and C5 (one tile per SM)
for c in C4 C5; do
python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 4736 > gpurun_out/${c}_small.json 2> gpurun_out/${c}_small.err && python -c "
import json;d=json.loads(open('gpurun_out/${c}_small.json').read().strip().splitlines()[-1]);print('$c small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_${c}_b4736 -f python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 4736 > gpurun_out/ncu_${c}.log 2>&1
tail -1 gpurun_out/ncu_${c}.log | cut -c1-120
done
