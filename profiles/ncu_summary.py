#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md and bench.py cite.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [substring of metric names to also print]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    # header entries look like "SECTION.metric" or plain "metric": index by the trailing metric name
    def short(h):
        parts = h.split(".")
        for i, p in enumerate(parts):
            if "__" in p:
                return ".".join(parts[i:])
        return h
    names = [short(h) for h in hdr]
    for r in data:
        d = {}
        for n, v, u in zip(names, r, units):
            d.setdefault(n, (v, u))
        print("== %s  grid %s block %s" % (d.get("Kernel Name", ("?",))[0][:70], d.get("Grid Size", ("?",))[0], d.get("Block Size", ("?",))[0]))
        for k in KEYS:
            if k in d:
                print("   %-82s %s %s" % (k, d[k][0], d[k][1]))
        for e in extra:
            for n in sorted(d):
                if e in n and n not in KEYS:
                    print("   %-82s %s %s" % (n, d[n][0], d[n][1]))


if __name__ == "__main__":
    main()
