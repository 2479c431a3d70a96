"""Summarise an `ncu --page raw --csv` export: one block per profiled launch with the metrics DESIGN.md / bench.py cite.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv
"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("----")
        print(f"{'Kernel Name':72s} {d.get('Kernel Name', '')}")
        for k in WANT:
            if k in d:
                print(f"{k:72s} {d[k]} {unit.get(k, '')}")
        st = [(float(v.replace(',', '')), k) for k, v in d.items()
              if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v]
        for v, k in sorted(st, reverse=True)[:7]:
            print(f"   stall {v:6.2f} {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")


if __name__ == "__main__":
    main(sys.argv[1])
