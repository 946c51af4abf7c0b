"""Print the key metrics of an .ncu-rep (one line per kernel) — used to build profiles/*.md."""
import csv, subprocess, sys
KEYS = [
 ("time_us", "gpu__time_duration.sum"),
 ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
 ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
 ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
 ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
 ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
 ("regs", "launch__registers_per_thread"),
 ("inst", "smsp__inst_executed.sum"),
 ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
 ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
 ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
 ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
 ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
 ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
 ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
 ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
 ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
 ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
 ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
 ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
 ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
 ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
 ("stall_not_sel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
 ("thr_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]][:70], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        for name, key in KEYS:
            if key in idx:
                v = r[idx[key]]
                try: v = f"{float(v):.4g}"
                except ValueError: pass
                print(f"   {name:18s} {v:>12s} {units[idx[key]]}")
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
