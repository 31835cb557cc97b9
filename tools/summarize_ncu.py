#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, without a GPU) into profiles/<name>.txt.

    python tools/summarize_ncu.py gpurun_out/prof_r1.ncu-rep profiles/ncu_full_r1.txt
"""
import csv, io, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# summary of {rep} (ncu --set full --clock-control none; per-launch values, cold cache, serialised)"]
    seen = {}
    for r in data:
        name = r[ki]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue  # first launch of each kernel is representative; the rest repeat it
        lines.append(f"\n== {name[:150]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"   {m:78s} {r[i]:>20s} {units[i]}")
    lines.append("\n# launches per kernel in the report: " + ", ".join(f"{v}x {k[:60]}" for k, v in seen.items()))
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"wrote {out}")


if __name__ == "__main__":
    main()
