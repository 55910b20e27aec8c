"""Condense an .ncu-rep (ncu --set full) into the per-kernel table kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.csv
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m, _ in METRICS if m in hdr}
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [f"{name}" for m, name in METRICS if m in idx])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.split("(")[0].replace("void ", "").replace("fbdev::", "")
        vals = []
        for m, _ in METRICS:
            if m not in idx:
                continue
            v = r[idx[m]]
            try:
                f = float(v)
                unit = units[idx[m]]
                if "dram" in m:
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
                if m == "gpu__time_duration.sum":
                    f *= {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(v)
        out.writerow([name] + vals)


if __name__ == "__main__":
    main(sys.argv[1])
