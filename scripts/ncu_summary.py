"""Summarise an .ncu-rep (read here, no GPU): python scripts/ncu_summary.py REP [regex] -> key metrics per launch."""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        print(f"== {name[:100]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:62s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
