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
    # per-kernel DRAM traffic (bytes per launch, last launch of each kernel) for bench.py's roofline.traffic
    import json, os
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    slot = {"k_lstm_fwd_tc": "lstm_fwd", "k_lstm_bwd_tc": "lstm_bwd", "k_wgrad_tc": "wgrad", "k_merw_walk": "merw_walk",
            "k_lstm_fwd_tc3": "lstm_fwd_tc3", "k_lstm_bwd_tc3": "lstm_bwd_tc3"}
    tpath = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).split("::")[-1]
        name = re.sub(r"<.*", "", name).replace("void ", "").strip()     # template instances: k_lstm_bwd_tc<1>
        if name in slot:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_read.sum"]], 1)
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_write.sum"]], 1)
            traffic[slot[name]] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "source": os.path.basename(rep)}
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
