"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export: a fixed list of metrics per kernel launch."""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "launch__shared_mem_per_block_static", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[col["Kernel Name"]].split("(")[0])
        for m in METRICS:
            if m in col:
                print(f"   {m:72s} {r[col[m]]:>16s} {units[col[m]]}")


if __name__ == "__main__":
    main(sys.argv[1])
