#!/usr/bin/env python
"""Hot source lines of one kernel from an ncu report (needs -lineinfo and --import-source on).
  python tools/ncu_hot.py gpurun_out/prof.ncu-rep k_raster_tiles [launch-index] [top-n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
tables, cur, path = [], None, ""
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "File Path":
        path = row[1]
    elif row and row[0] == "Function Name":
        cur = {"name": row[1], "hdr": None, "rows": [], "path": path}
        if path.endswith((".cuh", ".cu")) and "pcr" in path:
            tables.append(cur)
    elif cur is not None and row and row[0] == "Line No":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None and len(row) > 8 and row[0].isdigit() and row[2] == "-":
        cur["rows"].append(row)
print(f"{len(tables)} launch tables")
t = tables[which]
h = t["hdr"]
ie, sm = h.index("Instructions Executed"), h.index("# Samples")
body = t["rows"]
tot_i = sum(int(r[ie] or 0) for r in body); tot_s = sum(int(r[sm] or 0) for r in body)
print(f"{t['name'][:70]}  instr {tot_i} samples {tot_s}")
for r in sorted(body, key=lambda r: -int(r[sm] or 0))[:topn]:
    print(f"{100*int(r[sm] or 0)/max(tot_s,1):5.1f}% samp {100*int(r[ie] or 0)/max(tot_i,1):5.1f}% instr | L{r[0]:>4} {r[1].strip()[:120]}")
