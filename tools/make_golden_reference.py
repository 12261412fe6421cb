"""Golden vectors from the REFERENCE'S OWN generated objectives (run in the build container).

baseline/introduction-single-stock.cpp and baseline/multiple-substocks.cpp -- the reference's
checked-in output of g3_to_tmb() -- and inttest/codegeneration/ling.cpp are compiled unmodified against oracle/tmb_shim
(make -C oracle ref) and evaluated on our C1 / C2 model data and on seeded parameter vectors.
The results are stored in tests/golden/reference_c{1,2}.npz so that the parity tests can compare
the oracle AND the CUDA path with the reference's code on machines where /root/reference does not
exist (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gadget3_b200.configs import CONFIGS, parameter_batch  # noqa: E402
from oracle import ref_lib  # noqa: E402

N_VECTORS = 24
REPORT_TIMES = [0, 1, 2, 3, 4, 66, 67, 134, 135]

if __name__ == "__main__":
    ref_lib.build()
    for name in ("C1", "C2", "LING"):
        model, p = CONFIGS[name]()
        extras = ref_lib.ling_extras(p) if name == "LING" else {}
        ref = ref_lib.RefModel(name, model, p, **extras)
        full = np.tile(p.values(), (N_VECTORS, 1))
        full[1:, p.optimised_indices()] = parameter_batch(p, N_VECTORS - 1, seed=20261019)
        full[N_VECTORS - 1, p.index("retro_years")] = 2          # a shortened run
        out = {"par": full, "nll": ref.eval_batch(full), "switch": np.array(list(p.switch))}
        for vi in (0, 5):
            ref.eval(full[vi], reporting=True)
            out[f"v{vi}__nll"] = ref.report("nll")
            for cn, unit in zip(model.comp_names, model.comp_units):
                out[f"v{vi}__nll_{cn}__{unit}"] = ref.report(f"nll_{cn}__{unit}")
                mm = ref.report(f"{cn}_model__{unit}")
                if mm is not None and name != "LING":      # (ling's report = FALSE components keep no history)
                    out[f"v{vi}__{cn}_model__{unit}"] = mm
                prm = ref.report(f"{cn}_model__params")
                if prm is not None:
                    out[f"v{vi}__{cn}_model__params"] = prm
            out[f"v{vi}__nll_understocking__wgt"] = ref.report("nll_understocking__wgt")
            for sn in model.stock_names:
                for what in ("num", "wgt"):
                    a = ref.report(f"dstart_{sn}__{what}")
                    if a is not None:           # (the ling model has no g3a_report_detail)
                        out[f"v{vi}__dstart_{sn}__{what}"] = a[..., REPORT_TIMES]
                a = ref.report(f"detail_{sn}__renewalnum")
                if a is not None:
                    out[f"v{vi}__detail_{sn}__renewalnum"] = a[..., REPORT_TIMES]
            for pred, prey, _ in model.pp_info:
                for key in (f"detail_{prey}_{pred}__suit", f"detail_{prey}_{pred}__cons", f"detail_{prey}__predby_{pred}"):
                    a = ref.report(key)
                    if a is not None:
                        out[f"v{vi}__{key}"] = a[..., REPORT_TIMES]
                out[f"v{vi}__suit_{prey}_{pred}__report"] = ref.report(f"suit_{prey}_{pred}__report")
        out["report_times"] = np.array(REPORT_TIMES)
        path = os.path.join(ROOT, "tests", "golden", f"reference_{name.lower()}.npz")
        out = {k: v for k, v in out.items() if v is not None}      # (ling.cpp REPORT()s only nll_understocking__wgt)
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path), "bytes; nll[:4]", out["nll"][:4])
