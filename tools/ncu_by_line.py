"""Executed warp instructions per CUDA source line for the first kernel of an .ncu-rep, by joining ncu's SASS
page (per-instruction counts, in address order) with `nvdisasm -g` line info of the same kernel:
python tools/ncu_by_line.py rep.ncu-rep lib.so [top]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kname = rows[0][1]
hdr = rows[1]; iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
sass = [(r[iS].strip(), int(r[iE] or 0), int(r[iW] or 0)) for r in rows[2:] if len(r) > iE]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
base = re.sub(r"^void ", "", kname).split("<")[0].split("::")[-1]
targs = re.findall(r"\((?:int|bool)\)(\d+)", kname)
best = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, lines, line = None, {}, None
    for l in dis.split("\n"):
        m = re.match(r"\.text\.(\S+):", l)
        if m: cur = m.group(1); lines[cur] = []; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines[cur].append(line)
    for k, v in lines.items():
        if base in k and len(v) == len(sass):
            best = v
if best is None:
    sys.exit(f"no function named like {base} with {len(sass)} instructions found")
by, st = collections.Counter(), collections.Counter()
for (src, ex, ws), ln in zip(sass, best):
    by[ln] += ex; st[ln] += ws
tot = sum(by.values())
print(f"{kname}: {tot} warp instructions")
srcs = {}
for (f, n), ex in by.most_common(top):
    if f not in srcs:
        p = [q for q in glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f))]
        srcs[f] = open(p[0]).read().split("\n") if p else []
    text = srcs[f][n - 1].strip()[:90] if srcs[f] and n <= len(srcs[f]) else ""
    print(f"{100 * ex / tot:5.1f}%  stall {100 * st[(f, n)] / max(1, sum(st.values())):5.1f}%  {f}:{n:<4d} {text}")
