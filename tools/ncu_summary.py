"""Key metrics of every kernel in an .ncu-rep (from `ncu --set full`) as markdown.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.sum', 'sm__cycles_elapsed.max',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ['# %s' % rep, '']
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        out.append('## %s  (id %s)' % (r[col['Kernel Name']], r[col['ID']]))
        out.append('')
        out.append('| metric | value | unit |')
        out.append('|---|---:|---|')
        for k in KEYS:
            if k in col:
                out.append('| %s | %s | %s |' % (k, r[col[k]], units[col[k]]))
        tens = [h for h in hdr if 'tensor' in h and '.avg.pct_of_peak_sustained_active' in h and h not in KEYS]
        for k in tens:
            out.append('| %s | %s | %s |' % (k, r[col[k]], units[col[k]]))
        out.append('')
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
