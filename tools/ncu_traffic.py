#!/usr/bin/env python
"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in an
ncu --set full report and write profiles/traffic.json keyed by bench.py's kernel names.
usage: python tools/ncu_traffic.py gpurun_out/<tag>/prof.ncu-rep [more.ncu-rep ...]"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def bench_name(kname, order):
    """bench.py's kernel name for an ncu kernel name (launch order within one step for the repeated kernels)."""
    if "conv_in_planes" in kname: return "conv_in_planes"
    if "conv_in_tc" in kname: return "conv_in_tc"
    if "tsdf_elements" in kname: return "conv_in:elements"
    if "xz_finish" in kname: return "xz_finish"
    if "nchw_to_tall" in kname: return "nchw_to_tall:pre"
    if "scene_argmax" in kname: return "scene_argmax"
    if "pool_tall" in kname:
        order["pool"] = order.get("pool", 0) + 1
        return "maxpool:p0" if order["pool"] % 2 == 1 else "maxpool:p1"
    if "conv_tall_persistent" in kname:
        names = ["conv3x3:d0c1", "conv3x3:d0c2", "conv3x3:d1c1", "conv3x3:d1c2", "conv3x3:d2c1", "conv3x3:d2c2", "convT:u0", "conv3x3:u0c1",
                 "conv3x3:u0c2", "convT:u1", "conv3x3:u1c1", "conv3x3:u1c2+final"]
        i = order.get("conv", 0)
        order["conv"] = i + 1
        return names[i % 12]
    if "decode_points" in kname: return "decode_points:grasp+tsdf"
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
