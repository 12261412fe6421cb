"""Instruction counts and stall samples of a kernel by ENCLOSING source line of the kernel file (inlined callees folded in).
Usage: python tools/ncu_lines.py REP KERNEL_SUBSTR KERNEL_FILE EVALS CELLSTEPS_PER_EVAL [TOP]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, pat, kfile, evals, cs = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
top = int(sys.argv[6]) if len(sys.argv) > 6 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "gadget3_b200/csrc/libg3cuda.so")], cwd=td, capture_output=True)
dis = None
for cb in sorted(glob.glob(td + "/*.cubin")):
    d = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout.split("\n")
    if any(l.startswith(".text.") and pat in l for l in d):
        dis = d
        break
assert dis is not None, "kernel not found in any cubin of libg3cuda.so"
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and pat in l][0]
lastk, amap = None, {}
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln = m.group(1).split("/")[-1], int(m.group(2))
        m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        hit = [int(n) for ff, n in m2 if ff.endswith(kfile)]
        if f == kfile and not hit:
            lastk = ln
        elif hit:
            lastk = hit[-1]         # outermost frame inside the kernel file
        elif not m2:
            lastk = (f, ln)          # a non-inlined function of another file
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = lastk
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; base = int(rows[2][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byk, byop, bys, ti, ts = collections.Counter(), collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter), 0, 0
for rr in rows[2:]:
    off = int(rr[0], 16) - base
    ie = int(rr[ix["Instructions Executed"]]); ti += ie
    k = amap.get(off)
    byk[k] += ie
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", rr[1].strip())
    byop[k][m.group(2) if m else "?"] += ie
    for s in stalls:
        v = int(rr[ix[s]] or 0)
        bys[k][s[6:]] += v; ts += v
srcl = open(os.path.join(root, "gadget3_b200/csrc", kfile)).read().split("\n")
per = ti * 32 / evals / cs
print(f"thread instructions per cell-step: {per:.1f}; stall samples {ts}")
tot_s = collections.Counter()
for k in bys:
    tot_s.update(bys[k])
print("stalls overall: " + ", ".join(f"{s} {v / ts * 100:.1f}%" for s, v in tot_s.most_common(8)))
order = sorted(byk, key=(lambda k: -byk[k]) if os.environ.get("BY_INST") else (lambda k: -(sum(bys[k].values()))))
for k in order[:top]:
    v = byk[k]
    text = srcl[k - 1].strip()[:80] if isinstance(k, int) else str(k)
    ops = ", ".join(f"{o} {c / ti * per:.0f}" for o, c in byop[k].most_common(3))
    st = ", ".join(f"{s} {c / ts * 100:.1f}" for s, c in bys[k].most_common(3))
    print(f"{str(k):>6s} inst {v / ti * 100:5.1f}% ({v / ti * per:6.1f}) stall {sum(bys[k].values()) / ts * 100:5.1f}% [{st}] [{ops}]  {text}")
