# End-of-round check on a fresh box, as the driver does it: smoke(), the GPU tests, the default bench line, the reference arm.
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
(time python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err); tail -3 gpurun_out/bench_r02_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_default.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_spot_check']['max_rel_err_nll'], d.get('cpu_baseline',{}).get('value'))
for e in d.get('extra_configs',[]):
    print(e['config']['workload'][:30], e['value'], e['e2e']['value'], e['roofline']['frac'], e['parity_spot_check']['max_rel_err_nll'], e['parity_spot_check'].get('within_contract'))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err; tail -c 300 gpurun_out/bench_r02_reference_arm.json
