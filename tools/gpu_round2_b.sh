set -x
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -40
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_v1_C2_b4736 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel 4 --batch 4736 > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_small.log
