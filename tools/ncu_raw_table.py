"""Selected columns of an ncu --set full report as CSV (one row per profiled launch): python tools/ncu_raw_table.py rep.ncu-rep > out.csv"""
import csv, subprocess, sys
WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
stall = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
w = csv.writer(sys.stdout)
w.writerow(WANT + ['top_stalls_warps_per_issue'])
for r in rows[2:]:
    vals = []
    for c in WANT:
        v = r[hdr.index(c)] if c in hdr else ''
        vals.append(v[:100] if c == 'Kernel Name' else v)
    st = sorted([(float(r[i].replace(',', '')), hdr[i][34:-23]) for i in stall if r[i] not in ('', 'n/a')], reverse=True)[:5]
    w.writerow(vals + ['; '.join('%s %.2f' % (n, v) for v, n in st)])
