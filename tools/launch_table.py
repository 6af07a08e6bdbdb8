"""Condenses an `ncu --csv --metrics gpu__time_duration.sum` log into one line per launch:
python tools/launch_table.py gpurun_out/launches.csv [out.txt]"""
import csv, io, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; iK, iG, iB, iM, iV = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value"))
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum": continue
    k = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("nvpyr::", "")
    if not any(t in k for t in ("Kernel", "kernel")) or "at::" in k or "elementwise" in k: continue
    print(f"{float(r[iV].replace(',', '')) / 1e3:9.2f} us  {k:50s} grid {r[iG]:14s} block {r[iB]}", file=out)
