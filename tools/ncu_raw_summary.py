"""Print the headline raw metrics of an ncu report (per captured launch)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if name_i is not None:
        print("kernels:", sorted({r[name_i][:90] for r in rows[2:]}))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {rows[1][i]:>12s}  {[r[i] for r in rows[2:]]}")


if __name__ == "__main__":
    main(sys.argv[1])
