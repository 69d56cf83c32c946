#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into the handful of numbers we track.
usage: python tools/ncu_summary.py gpurun_out/<tag>/prof_x.ncu-rep [> profiles/<name>.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput%"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "math_pipe_throttle", "not_selected", "wait",
          "dispatch_stall", "lg_throttle", "no_instruction", "branch_resolving", "membar", "sleeping", "selected"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("=" * 100)
        print(r[col["Kernel Name"]][:160])
        for k, label in KEYS:
            if k in col:
                print(f"  {label:<22}{r[col[k]]:>16} {units[col[k]]}")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in col:
                try:
                    st.append((float(r[col[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls (warps per issue-active cycle): " + ", ".join(f"{s}={v:.2f}" for v, s in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
