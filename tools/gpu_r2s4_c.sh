python - <<'PY'
import numpy as np
from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import g3_cuda_adfun
m, p = CONFIGS["SIX_FLEETS"]()
fn = g3_cuda_adfun(m, p, device=0)
print("SIX_FLEETS", fn.handle.tile_info())
try:
    fn.handle.set_kernel(3)
except Exception as e:
    print("why:", e)
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
