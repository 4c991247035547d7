#!/usr/bin/env python
"""Text summary of an ncu capture for profiles/: the metrics the B200 profiling recipe names, per kernel launch.

    python scripts/ncu_summary.py gpurun_out/prof_tc_r2_final.ncu-rep > profiles/glm_tc_r2_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main() -> None:
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        vals = {}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                vals[key] = (r[i], units[i])
                print(f"{key:75s} {r[i]:>18s} {units[i]}")
        try:
            t = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
            unit = vals["gpu__time_duration.sum"][1]
            seconds = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
            rb = float(vals["dram__bytes_read.sum"][0].replace(",", ""))
            rb *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(vals["dram__bytes_read.sum"][1], 1)
            print(f"{'derived: DRAM read bandwidth':75s} {rb / seconds / 1e12:18.3f} TB/s")
        except Exception:
            pass


if __name__ == "__main__":
    main()
