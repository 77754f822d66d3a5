"""Print the key roofline metrics of every kernel in an ncu report.
usage: python scripts/ncu_summary.py gpurun_out/<tag>/full.ncu-rep [out.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    print('----')
    rec = {}
    for w in want:
        if w in idx:
            print(f"  {w} [{units[idx[w]]}] = {r[idx[w]][:100]}")
            rec[w] = r[idx[w]]
    out.append(rec)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w", newline="") as f:
        wr = csv.DictWriter(f, fieldnames=[w for w in want if w in idx]); wr.writeheader(); wr.writerows(out)
