#!/usr/bin/env python
"""Compact text summary of an Nsight Compute report (run HERE, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.txt
One block per profiled launch: duration, DRAM bytes (the `traffic` of bench.py's roofline object), DRAM / L2 /
tensor-pipe / SM utilisation, occupancy facts."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; cold caches, serialised)")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']][:100]}")
        for m, label in METRICS:
            if m in col and r[col[m]] != "":
                print(f"   {label:32s} {r[col[m]]} {units[col[m]]}")
        try:
            rd, wr = float(r[col["dram__bytes_read.sum"]].replace(",", "")), float(r[col["dram__bytes_write.sum"]].replace(",", ""))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * scale[units[col["dram__bytes_read.sum"]]] + wr * scale[units[col["dram__bytes_write.sum"]]]
            print(f"   {'traffic (dram read + write)':32s} {tot / 1e6:.3f} MB")
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
