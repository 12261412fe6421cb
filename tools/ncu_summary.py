"""Summarise an .ncu-rep (raw page) and, with --lines, aggregate SASS metrics by source line.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines KERNEL_MANGLED_SUBSTR]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, r):
    if h in keys:
        print(f"{h:70s} {v} {u}")
for h, u, v in zip(hdr, units, r):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v) > 0.05:
        print(f"{h[34:-23]:40s} {float(v):.3f}")
if len(sys.argv) > 3 and sys.argv[2] == "--lines":
    import glob, os, tempfile
    pat = sys.argv[3]
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gadget3_b200", "csrc", "libg3cuda.so")
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = glob.glob(os.path.join(td, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and pat in l][0]
    cur, amap = None, {}
    for l in dis[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            amap[int(m.group(1), 16)] = cur
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(rows[2][0], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    ops = collections.Counter()
    ti = ts = 0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for rr in rows[2:]:
        off = int(rr[0], 16) - base
        a = agg[amap.get(off)]
        ie, sa = int(rr[ix["Instructions Executed"]]), int(rr[ix["# Samples"]])
        a[0] += ie; a[1] += sa; a[2] += 1
        for k in stalls:
            a[3][k] += int(rr[ix[k]])
        ti += ie; ts += sa
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", rr[1].strip())
        ops[m.group(2) if m else "?"] += ie
    print("total warp instructions", ti, "samples", ts, "SASS lines", len(rows) - 2)
    for srcl, (ie, sa, n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(srcl, f"inst {ie / ti * 100:5.1f}% samp {sa / ts * 100:5.1f}% sass {n}", dict(c.most_common(3)))
    print("opcode mix:", ", ".join(f"{o} {c / ti * 100:.1f}%" for o, c in ops.most_common(24)))
