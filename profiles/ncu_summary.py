"""Key metrics of every kernel in an `ncu --set full` report, as a markdown table.
usage: python profiles/ncu_summary.py report.ncu-rep [more reports...]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % peak"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print(f"\n### {rep.split('/')[-1]}\n")
    kernels = rows[2:]
    print("| metric | " + " | ".join(f"`{r[ki].split('(')[0][:34]}`" for r in kernels) + " |")
    print("|---|" + "---:|" * len(kernels))
    for key, label in WANT:
        if key not in h:
            continue
        i = h.index(key)
        vals = []
        for r in kernels:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.3g} {units[i]}" if abs(v) < 1e4 else f"{v:.4g} {units[i]}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} | " + " | ".join(vals) + " |")
