"""Compact one-line-per-launch table from an `ncu --page raw --csv` export.
    python tools/ncu_table.py gpurun_out/x_raw.csv [out.md]
"""
import csv
import sys

COLS = [('us', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'), ('regs', 'launch__registers_per_thread'),
        ('dramR_MB', 'dram__bytes_read.sum'), ('dramW_MB', 'dram__bytes_write.sum'),
        ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('lts%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('l1%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('warps%', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('Minst', 'smsp__inst_executed.sum'),
        ('fma%', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
        ('lsu%', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
        ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('st_long', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
        ('st_short', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
        ('st_bar', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
        ('st_wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
        ('st_math', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio'),
        ('st_mio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio'),
        ('st_lg', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio')]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'inst': 1e-6}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = ['| kernel | ' + ' | '.join(c for c, _ in COLS) + ' |', '|---|' + '---:|' * len(COLS)]
    for r in rows[2:]:
        cells = []
        for name, k in COLS:
            if k not in col or r[col[k]] in ('', 'n/a'):
                cells.append('-')
                continue
            v, u = r[col[k]].replace(',', ''), units[col[k]]
            try:
                f = float(v) * (SCALE.get(u, 1.0) if not name.startswith('st_') else 1.0)
                cells.append('%.0f' % f if name in ('grid', 'regs') else '%.2f' % f)
            except ValueError:
                cells.append(v)
        out.append('| %s | %s |' % (r[col['Kernel Name']].split('(')[0][:40], ' | '.join(cells)))
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write('# %s (ncu --set full, --clock-control none)\n\n' % sys.argv[1] + text + '\n')


if __name__ == '__main__':
    main()
