"""ncu --set full report -> a small CSV of the numbers the roofline discussion uses (one row per profiled launch).

    python profiles/ncu_summary.py gpurun_out/r2/n23_m.ncu-rep > profiles/r2_ncu_n2_n3_sort_take.csv
"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k) if k in hdr else -1, name) for k, name in WANT]
    w = csv.writer(sys.stdout)
    w.writerow([name for _, name in idx])
    for r in rows[2:]:
        out = []
        for i, name in idx:
            if i < 0:
                out.append("")
                continue
            v = r[i]
            if name == "kernel":
                v = v.split("(")[0].replace("void ", "")
            elif name in ("dram_read_MB", "dram_write_MB"):
                u = units[i]
                f = float(v.replace(",", "")) if v else 0.0
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                v = f"{f:.1f}"
            elif name == "duration_us":
                u = units[i]
                f = float(v.replace(",", "")) if v else 0.0
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                v = f"{f:.1f}"
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main()
