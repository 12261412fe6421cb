# GPU suite, then the default bench without the CPU baseline; prints the headline numbers of every config
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; tail -c 400 gpurun_out/bench_check.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_check.json").read().strip().splitlines()[-1])
print("C2", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity_spot_check"]["max_rel_err_nll"])
for e in d.get("extra_configs",[]): print(e["config"]["workload"][:3], e.get("value"), e.get("roofline",{}).get("frac"))
PY
