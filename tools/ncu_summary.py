"""Prints the key columns of an ncu report (raw page): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    name = hdr.index('Kernel Name')
    st = [i for i, h in enumerate(hdr) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print('==', r[name][:90])
        for w in WANT:
            if w in hdr:
                print('   %-70s %s' % (w, r[hdr.index(w)]))
        vals = sorted([(float(r[i].replace(',', '')), hdr[i][34:-23]) for i in st if r[i] not in ('', 'n/a')], reverse=True)[:7]
        print('   stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in vals))


if __name__ == '__main__':
    main()
