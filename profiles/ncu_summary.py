"""Print the metrics we track from an .ncu-rep (reads the raw page via `ncu -i ... --csv`)."""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg.per_second',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio','smsp__inst_executed.sum']
def main(path, grep=None):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')][:70])
        for i,h in enumerate(hdr):
            if (h in WANT) or (grep and grep in h):
                print(f'  {h:85s} {r[i]:>16s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
