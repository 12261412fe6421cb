(time python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err); tail -3 gpurun_out/bench_r02_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_default.json'))
print('C2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_spot_check'], d.get('cpu_baseline',{}).get('value'))
for e in d.get('extra_configs',[]):
    print(e['config']['workload'][:30], e['value'], e['e2e']['value'], e['roofline']['frac'], e['roofline']['kernel'][:30], e['parity_spot_check'])
PY
