"""Summaries of ncu exports (made on the GPU box with `ncu -i X.ncu-rep --page raw --csv` and
`--page source --csv --print-source cuda,sass`): headline metrics and stall samples / instructions /
shared-memory wavefronts per CUDA source line.   python scripts/ncu_summarise.py raw.csv src.csv [top]"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__sass_branch_targets_threads_uniform.pct', 'smsp__sass_average_branch_targets_threads_uniform.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
print("# headline metrics (ncu --set full, one launch)")
for i, h in enumerate(hdr):
    if h in WANT:
        print("%-82s %-14s %s" % (h, units[i], vals[i][:110]))
rows = list(csv.reader(open(src)))
cur, agg, h2 = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h2 = r; continue
    if r[0].isdigit() and h2:
        try:
            wi = h2.index("L1 Wavefronts Shared")
            agg.append((cur, int(r[0]), r[1], int(r[4]), int(r[7]), int(r[wi]) if r[wi].isdigit() else 0))
        except (ValueError, IndexError):
            pass
ts, ti, tw = sum(a[3] for a in agg), sum(a[4] for a in agg), max(1, sum(a[5] for a in agg))
print("\n# per CUDA source line: %d stall samples, %d warp instructions, %d shared-memory wavefronts" % (ts, ti, tw))
print("%-22s %5s %8s %8s %8s  %s" % ("file", "line", "samples", "inst", "smem wf", "source"))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-22s %5d %7.2f%% %7.2f%% %7.2f%%  %s" % (a[0][:22], a[1], 100 * a[3] / ts, 100 * a[4] / ti, 100 * a[5] / tw, a[2].strip()[:96]))
