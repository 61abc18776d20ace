"""Condense `ncu -i X.ncu-rep --page raw --csv` into one block per kernel launch (for profiles/)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ki][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"    {k:70s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("    top stall reasons (share of samples): " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in stalls[:6]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
