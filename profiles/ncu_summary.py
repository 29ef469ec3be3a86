#!/usr/bin/env python
"""Text summary of an ncu report for profiles/ (the .ncu-rep files themselves stay in gpurun_out/, which is scratch):
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep "header line: the command that was profiled" > profiles/rNN_x_ncu.txt
Keeps the metrics the roofline arguments of DESIGN.md / bench.py rest on."""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active", "sm__pipe_fp64_cycles_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_", "l1tex__data_pipe_lsu_wavefronts", "smsp__average_warps_issue_stalled", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct",
        "sm__inst_executed_pipe_tc", "utcimma", "smsp__inst_executed_op_shared", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct")


def main():
    rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# " + header)
    for li, vals in enumerate(rows[2:]):
        d = dict(zip(hdr, vals))
        print(f"## launch {li}")
        print(d.get("Kernel Name", ""), d.get("Grid Size", ""), "x", d.get("Block Size", ""))
        for h, u, v in zip(hdr, units, vals):
            if any(k in h for k in KEEP) and "pcsamp" not in h and v not in ("", "n/a"):
                print(f"{h:120s} {u:>16s} {v:>20s}")


if __name__ == "__main__":
    main()
