"""Markdown summary of the kernels in `ncu --page raw --csv` exports (profiles/ncu_r02/*_raw.csv): the counters DESIGN.md and
bench.py quote.   python tools/ncu_summary.py profiles/ncu_r02/*_raw.csv > profiles/ncu_r02/SUMMARY.md"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration"),
        ("launch__registers_per_thread", "registers / thread"),
        ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "FP64+DMMA pipe (sm__pipe_shared_cycles_active)"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "  of which FP64 (DFMA/DMUL/DADD)"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "  of which DMMA"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch / issue"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (of peak)"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (of peak)"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("smsp__inst_executed.sum", "warp instructions executed")]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        h, u = rows[0], rows[1]
        seen = {}
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            if r[h.index("gpu__time_duration.sum")] in ("", "nan", "-nan") or "nan" in r[h.index("smsp__inst_executed.sum")]:
                continue
            grid = r[h.index("Grid Size")] if "Grid Size" in h else ""
            seen[(name, grid)] = seen.get((name, grid), 0) + 1
            if seen[(name, grid)] > 1:          # repeated launches of one kernel with one grid: the first stands for all
                continue
            print(f"## `{name}`  grid {grid}  ({path.split('/')[-1]})\n")
            print("| counter | value |\n|---|---|")
            for k, label in KEYS:
                if k in h:
                    i = h.index(k)
                    val = r[i]
                    try:
                        val = f"{float(val):.4g}"
                    except ValueError:
                        pass
                    print(f"| {label} (`{k}`) | {val} {u[i]} |")
            print()


if __name__ == "__main__":
    main()
