#!/usr/bin/env python
"""profiles/ncu_counters.json from an ncu capture of ONE device-resident step (bench.py --profile-one-step under
`ncu --profile-from-start off --set full`): per kernel the DRAM bytes and warp instructions of its launches, per frame —
what bench.py reports as roofline.traffic and roofline.issue.

  python tools/ncu_counters.py gpurun_out/r02_step.ncu-rep H 32 profiles/ncu_counters.json "label of the capture"
"""
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, workload, frames, dst = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    label = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    col = {m: h.index(m) for m in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "gpu__time_duration.sum")}
    units = rows[1]

    def val(r, m):
        v = float(r[col[m]].replace(",", ""))
        u = units[col[m]]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "inst": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    kernels, total_inst, total_dram = {}, 0.0, 0.0
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        k = kernels.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "warp_instructions": 0.0, "us_under_ncu": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        k["warp_instructions"] += val(r, "smsp__inst_executed.sum")
        k["us_under_ncu"] += val(r, "gpu__time_duration.sum")
    for name, k in kernels.items():
        # a step issues exactly ONE K0 (k_stats + the serial mean, for a later step); a capture whose step had no hint yet also
        # holds that step's own in-line K0: count one launch of each
        share = 1.0 / k["launches"] if name in ("k_stats", "k_mean_sequential") and k["launches"] > 1 else 1.0
        total_inst += k["warp_instructions"] * share
        total_dram += k["dram_bytes"] * share
        k["dram_bytes_per_frame_per_launch"] = k["dram_bytes"] / k["launches"] / frames
        k["warp_instructions_per_frame"] = k["warp_instructions"] / frames
    data = json.load(open(dst)) if os.path.isfile(dst) else {}
    data[workload] = {"capture": label, "frames_per_step": frames, "warp_instructions_per_frame": total_inst / frames,
                      "dram_bytes_per_frame": total_dram / frames, "kernels": kernels}
    json.dump(data, open(dst, "w"), indent=1)
    print(json.dumps({k: (round(v["warp_instructions"] / 1e6, 1), round(v["dram_bytes"] / 1e6, 1), round(v["us_under_ncu"], 1)) for k, v in kernels.items()}, indent=1))
    print("total warp instructions / step: %.1f M, DRAM bytes / step: %.1f MB" % (total_inst / 1e6, total_dram / 1e6))


if __name__ == "__main__":
    main()
