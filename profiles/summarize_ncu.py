"""Summarise an `ncu --page raw --csv` export: one block of key metrics per distinct kernel.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu.py raw.csv > profiles/<name>.md
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        name = r[name_i].split("(")[0]
        seen.setdefault(name, []).append(r)
    print(f"# ncu --set full summary ({path}); per-launch values of the first captured launch, n = launches captured\n")
    for name, rs in seen.items():
        r = rs[0]
        print(f"## {name}  (n={len(rs)})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', ''))
        wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', ''))
        u = units[hdr.index('dram__bytes_read.sum')]
        print(f"| traffic (read+write) | {rd + wr:.3f} | {u} |\n")


if __name__ == "__main__":
    main(sys.argv[1])
