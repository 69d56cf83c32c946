#!/usr/bin/env python
"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an
ncu --set full report and write profiles/traffic.json keyed by bench.py's kernel names.
usage: python tools/ncu_traffic.py gpurun_out/<tag>/prof.ncu-rep [more.ncu-rep ...]"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def bench_name(kname, order):
    if "conv_in_planes" in kname: return "conv_in_planes"
    if "conv_in_tc" in kname: return "conv_in_tc"
    if "yz_finish" in kname: return "yz_finish"
    if "xz_finish" in kname: return "xz_finish"
    if "nchw_to_tall" in kname: return "nchw_to_tall:pre"
    if "scene_argmax" in kname: return "scene_argmax"
    if "decode_points" in kname:
        order["dec"] = order.get("dec", 0) + 1
        return "decode_points:grasp" if order["dec"] % 2 == 1 else "decode_points:tsdf"
    return None


out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    order = {}
    for r in rows[2:]:
        name = bench_name(r[col["Kernel Name"]], order)
        rd = float(r[col["dram__bytes_read.sum"]]) * unit_scale(units[col["dram__bytes_read.sum"]])
        wr = float(r[col["dram__bytes_write.sum"]]) * unit_scale(units[col["dram__bytes_write.sum"]])
        if name and name not in out:
            out[name] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "source": os.path.basename(os.path.dirname(rep)) + "/" + os.path.basename(rep)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
