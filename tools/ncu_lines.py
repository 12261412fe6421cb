"""Instruction counts of a kernel by ENCLOSING source line of the kernel file (inlined callees folded in).
Usage: python tools/ncu_lines.py REP KERNEL_SUBSTR KERNEL_FILE EVALS CELLSTEPS_PER_EVAL"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, pat, kfile, evals, cs = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "gadget3_b200/csrc/libg3cuda.so")], cwd=td, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(td + "/*.cubin")[0]], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and pat in l][0]
lastk, amap, opmap = None, {}, {}
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln = m.group(1).split("/")[-1], int(m.group(2))
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        if f == kfile:
            lastk = ln
        elif m2 and kfile in m2.group(1):
            lastk = int(m2.group(2))
        elif not m2:
            lastk = (f, ln)          # a non-inlined function of another file
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = lastk
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; base = int(rows[2][0], 16)
byk, byop, ti = collections.Counter(), collections.defaultdict(collections.Counter), 0
for rr in rows[2:]:
    off = int(rr[0], 16) - base
    ie = int(rr[ix["Instructions Executed"]]); ti += ie
    k = amap.get(off)
    byk[k] += ie
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", rr[1].strip())
    byop[k][m.group(2) if m else "?"] += ie
srcl = open(os.path.join(root, "gadget3_b200/csrc", kfile)).read().split("\n")
per = ti * 32 / evals / cs
print(f"thread instructions per cell-step: {per:.1f}")
for k, v in byk.most_common(45):
    text = srcl[k - 1].strip()[:90] if isinstance(k, int) else str(k)
    ops = ", ".join(f"{o} {c / ti * per:.0f}" for o, c in byop[k].most_common(4))
    print(f"{str(k):>28s} {v / ti * 100:5.1f}% {v / ti * per:6.1f}  [{ops}]  {text}")
