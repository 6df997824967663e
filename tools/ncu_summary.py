"""Summary of one `ncu --page raw --csv` export: the metric set profiles/r*_ncu_*.txt keeps.  Usage: python tools/ncu_summary.py RAW.csv [particles]"""
import csv, sys

KEEP = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum launch__block_size launch__grid_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
lts__t_sectors_srcunit_tex_op_red.sum lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio""".split()

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print("kernel:", vals[col["Kernel Name"]])
out = {}
for k in KEEP:
    if k in col:
        out[k] = (vals[col[k]], units[col[k]])
        print(f"{k:90s} {vals[col[k]]:>18s} {units[col[k]]}")
if len(sys.argv) > 2:
    n = float(sys.argv[2])
    f = lambda k: float(out[k][0].replace(",", ""))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    b = sum(f(k) * scale[out[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(f"# DRAM traffic {b / n:.1f} B/particle; {f('smsp__inst_executed.sum') / n:.1f} warp-instructions/particle")
