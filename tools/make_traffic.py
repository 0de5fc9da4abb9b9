#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full captures of tools/profile_step.py (one 300-frame launch group each):
per kernel class the mean dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the geometry string
bench.py uses, together with the SHA-1 of the CUDA sources the captures were taken with (bench.py reports `traffic`
only while that still matches).

  python tools/make_traffic.py <commit> <geom_key>=<report.ncu-rep> [...]
  e.g. 640x480_L3_S1_B300_ppt128_frame=gpurun_out/r2h_main_raw.csv
Also writes profiles/<report>_summary.csv (the selected metrics of every captured launch) next to it."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def launches(rep):
    if rep.endswith(".csv"):  # the raw page exported on the GPU box (tools/ncu_round.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        d["_units"] = dict(zip(head, units))
        out.append(d)
    return out


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(val.replace(",", "")) * scale


def classify(name, grid, levels_px):
    if "k_ingest" in name:
        return "k_ingest"
    if "k_normals" in name:
        return "k_normals"
    if "k_compose" in name:
        return "k_compose"
    if "k_icp" in name:
        return None  # by launch order, see below
    return name.split("(")[0]


def main():
    import bench
    import ncu_summarize

    commit = sys.argv[1]
    captures = {}
    for arg in sys.argv[2:]:
        key, rep = arg.split("=", 1)
        levels = int(key.split("_L")[1].split("_")[0])
        iters = [10, 5, 4, 4][:levels]
        order = []  # k_icp launches of a group run coarse -> fine
        for l in range(levels - 1, -1, -1):
            order += [f"k_icp_L{l}"] * iters[l]
        per = {}
        icp_i = 0
        for d in launches(rep):
            name = d["Kernel Name"]
            cls = classify(name, d.get("Grid Size"), None)
            if cls is None:
                cls = order[icp_i] if icp_i < len(order) else "k_icp_extra"
                icp_i += 1
            u = d["_units"]
            b = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            per.setdefault(cls, []).append(b)
        name = os.path.splitext(os.path.basename(rep))[0].replace("_raw", "")
        summary = os.path.join("profiles", name + "_ncu_full_summary.csv")
        ncu_summarize.full(rep, os.path.join(ROOT, summary))
        captures[key] = {"source": summary + " (ncu --set full --clock-control none of tools/profile_step.py, one launch group)",
                         "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in per.items()},
                         "launches_captured": {k: len(v) for k, v in per.items()}}
    out = {"captured_at_commit": commit, "kernels_sha1": bench.kernels_sha1(),
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of a class",
           "captures": captures}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
