#!/usr/bin/env python3
"""ncu_summary.py — `ncu -i X.ncu-rep --page raw --csv` → the handful of metrics the roofline discussion uses, as a markdown table.

    python tools/ncu_summary.py gpurun_out/c2b_k2_pair_kernel_full.raw.csv "title" "command" > profiles/r02_....summary.md
"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__cluster_dim_y",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]


def main():
    path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(path, newline="")))
    header, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}\n\nCommand: `{command}`\n")
    for n, row in enumerate(data):
        name = row[header.index("Kernel Name")] if "Kernel Name" in header else "?"
        print(f"Launch {n}: `{name}`\n\n| metric | value | unit |\n|---|---|---|")
        for m in WANT:
            if m in header:
                i = header.index(m)
                print(f"| `{m}` | {row[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main()
