#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an ncu report (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; per = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if r[0].isdigit():
        try: s = int(r[4])
        except ValueError: s = 0
        ie = hdr.index("Instructions Executed")
        try: inst = int(r[ie])
        except ValueError: inst = 0
        key = (cur, int(r[0]))
        old = per.get(key, (0, 0, ""))
        per[key] = (old[0] + s, old[1] + inst, r[1].strip()[:100])
tot = sum(v[0] for v in per.values()); toti = sum(v[1] for v in per.values())
print(f"total samples {tot}, warp instructions {toti}")
for (f, l), (s, inst, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}% {s:6d} inst={100*inst/max(1,toti):5.1f}% {f}:{l}: {src}")
