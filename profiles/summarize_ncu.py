#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, one kernel captured with --set full) into the tracked summary files:

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<out_stem> "<title>" ["<command line>"]

writes <out_stem>_summary.md (headline metrics + stall reasons) and <out_stem>_details.csv (every raw
metric).  Reads the report with `ncu -i ... --page raw --csv`; needs no GPU."""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * mult.get(unit, 1)


def main():
    rep, stem, title = sys.argv[1], sys.argv[2], sys.argv[3]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(stem + "_details.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            w.writerow([h, u, v])
    get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = ["# " + title, ""]
    if cmd:
        out += ["Command: `%s`" % cmd, ""]
    out += ["| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in get:
            out.append("| %s | %s | %s |" % (k, get[k][0], get[k][1]))
    stalls = sorted(((float(v.replace(",", "") or 0), h) for h, (v, u) in get.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                     and not h.endswith("_not_issued")), reverse=True)
    for n, h in stalls[:8]:
        out.append("| %s | %d | samples |" % (h, n))
    if "dram__bytes_read.sum" in get and "dram__bytes_write.sum" in get:
        tot = to_bytes(*get["dram__bytes_read.sum"]) + to_bytes(*get["dram__bytes_write.sum"])
        out += ["", "DRAM traffic of this launch: %.1f MB (read + write)." % (tot / 1e6)]
    open(stem + "_summary.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
