#!/bin/bash
# GPU box: throughput of kernel build variants under gadget3_b200/csrc/variants/ (experiment helper).
for cfg in ${CFGS:-C2 C1}; do
  for v in "" $(ls gadget3_b200/csrc/variants/*.so 2>/dev/null); do
    r=$(G3CUDA_LIBRARY=$v python bench.py --config $cfg --steps 3 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4))")
    echo "$cfg ${v:-default} $r"
  done
done
