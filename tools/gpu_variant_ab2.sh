#!/bin/bash
# GPU box: C4 with the lean11 build (no migration / stock predators / other likelihoods), C5 with lean12 (no constant maturity / other likelihoods)
run() { G3CUDA_LIBRARY=$2 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4))"; }
echo "C4 default $(run C4 '')"; echo "C4 lean11 $(run C4 gadget3_b200/csrc/variants/libg3cuda_lean11.so)"
echo "C5 default $(run C5 '')"; echo "C5 lean12 $(run C5 gadget3_b200/csrc/variants/libg3cuda_lean12.so)"
echo "C2 lean11 $(run C2 gadget3_b200/csrc/variants/libg3cuda_lean11.so)"; echo "C2 lean12 $(run C2 gadget3_b200/csrc/variants/libg3cuda_lean12.so)"
