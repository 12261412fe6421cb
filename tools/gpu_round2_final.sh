# Final captures of round 2 (same commit): default bench line, reference arm, one --set full capture of the tile kernel
# (C2, one tile per SM), the launch list of a short bench run.
(time python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err); tail -3 gpurun_out/bench_r02_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_default.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_spot_check']['max_rel_err_nll'], d.get('cpu_baseline',{}).get('value'))
for e in d.get('extra_configs',[]):
    print(e['config']['workload'][:30], e['value'], e['e2e']['value'], e['roofline']['frac'], e['roofline']['kernel'][:30], e['parity_spot_check']['max_rel_err_nll'], e['parity_spot_check'].get('max_rel_err_grad'))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err; tail -c 400 gpurun_out/bench_r02_reference_arm.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 4736 > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err && python -c "
import json;d=json.loads(open('gpurun_out/plain_small.json').read().strip().splitlines()[-1]);print('small',d['value'],d['ms_per_step'])" &&
ncu --set full --clock-control none --import-source on -k regex:eval_tile -s 1 -c 1 -o gpurun_out/r02_tile_v8_C2_b4736 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --batch 4736 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/launches_r02.csv | cut -c1-200
