#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_tc.ncu-rep profiles/r01/ncu_tc_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "sm__cycles_elapsed.avg"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 else []
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep}: ncu --set full --clock-control none (values per launch)"]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        lines.append(f"== {name}  (launch id {r[idx['ID']]})")
        for k in KEYS:
            if k in idx:
                lines.append(f"   {k:75s} {r[idx[k]]:>18s} {units[idx[k]] if units else ''}")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            lines.append(f"   {'traffic = dram read + write (as reported units)':75s} {rd + wr:18.3f}")
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
