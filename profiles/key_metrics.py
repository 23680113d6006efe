"""Print the headline metrics of an ncu report: python profiles/key_metrics.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {vals[i]:>18s} {units[i]}")
    for i, h in enumerate(hdr):
        if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") and float(vals[i] or 0) > 3:
            print(f"  stall {h[len('smsp__warp_issue_stalled_'):-len('_per_warp_active.pct')]:40s} {float(vals[i]):8.1f} %")
    # warps stalled per issued instruction, by reason (sums to the average number of resident warps per issue)
    st = [(h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(vals[i] or 0)) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    tot = sum(v for _, v in st) or 1.0
    for name, v in sorted(st, key=lambda x: -x[1]):
        if v / tot > 0.02:
            print(f"  stalled warps per issue: {name:28s} {v:7.2f}  ({100 * v / tot:4.1f} %)")
