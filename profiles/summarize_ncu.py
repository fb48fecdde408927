#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small per-kernel summary that is committed under profiles/.

usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx   -> r01_xxx.md (+ .json)
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d["Kernel Name"], "block": d.get("Block Size"), "grid": d.get("Grid Size")}
        for k in KEYS:
            if k in d:
                e[k] = [d[k], units[hdr.index(k)]]
        res.append(e)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}` (read with `ncu -i ... --page raw --csv`)\n\n")
        for e in res:
            f.write(f"## {e['kernel']}\n\ngrid {e['grid']} block {e['block']}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in e:
                    f.write(f"| {k} | {e[k][0]} | {e[k][1]} |\n")
            f.write("\n")
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
