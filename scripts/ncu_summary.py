#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box) into a small text file for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name.txt ["note"]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep}", f"# {note}"]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        lines.append(f"kernel: {r[name_i]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"  {k:88s} {r[i]:>16s} {units[i]}")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', '')); ru = units[hdr.index('dram__bytes_read.sum')]
            wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', '')); wu = units[hdr.index('dram__bytes_write.sum')]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * mul[ru] + wr * mul[wu]
            dur = float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[hdr.index('gpu__time_duration.sum')]]
            lines.append(f"  derived: dram traffic {tot / 1e6:.1f} MB per launch, {tot / dur / 1e9:.0f} GB/s under ncu (cold cache, serialised)")
        except Exception as e:
            lines.append(f"  derived: n/a ({e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
