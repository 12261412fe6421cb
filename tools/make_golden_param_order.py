"""Extract the PARAMETER() declaration order of the reference's checked-in generated objectives
into tests/golden/ (run in the build container, where /root/reference exists).

The order is the column order of the `par` matrix at the C ABI (obj$env$last.par order,
R/run_tmb.R:715): g3_to_cuda() must reproduce it exactly.
"""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, out in (("introduction-single-stock", "c1"), ("multiple-substocks", "c2")):
    src = open(os.path.join(REF, "baseline", name + ".cpp")).read()
    pars = re.findall(r"^\s*PARAMETER\((\w+)\);", src, flags=re.M)
    with open(os.path.join(OUT, f"parameter_order_{out}.txt"), "w") as fh:
        fh.write(f"# PARAMETER() order of baseline/{name}.cpp (cpp-escaped names)\n")
        fh.write("\n".join(pars) + "\n")
    bounds = re.findall(r"// g3l_bounds_penalty for ([\w.]+);", src)
    with open(os.path.join(OUT, f"bounds_order_{out}.txt"), "w") as fh:
        fh.write(f"# run order of the g3l_bounds_penalty blocks in baseline/{name}.cpp (parameter switches)\n")
        fh.write("\n".join(bounds) + "\n")
    print(out, len(pars), len(bounds))
