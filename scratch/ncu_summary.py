"""Print the key metrics of every kernel in an .ncu-rep (read here, no GPU)."""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'sm__inst_executed.avg.per_cycle_elapsed', 'launch__grid_size', 'launch__block_size', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor', 'smsp__cycles_active.avg']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, u = r[0], r[1]
for v in r[2:]:
    print("==", v[h.index("Kernel Name")][:90])
    for i, k in enumerate(h):
        if k in KEYS:
            print(f"   {k} [{u[i]}] = {v[i]}")
    st = [(float(v[i] or 0), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i, k in enumerate(h)
          if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio')]
    print("   stalls:", ", ".join(f"{n}={x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
