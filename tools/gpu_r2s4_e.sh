# session 4, call e: prefetches in the slab-state paths (collect-vector conversion, next block's state rows)
for c in C4 C5 C2; do
python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', d['value'], d['parity_spot_check']['max_rel_err_nll'], d['parity_spot_check']['within_contract'])"
done
python tools/gpu_grad_warps.py C2 4096 15 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
