"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_summary.py gpurun_out/launches.csv [out.md]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(r['Metric Value'].replace(',', ''))
            unit = r['Metric Unit']
            v_us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
            rows.append((r['Kernel Name'].split('(')[0], v_us, r['Grid Size'], r['Block Size']))
    tot = collections.OrderedDict()
    for k, v, g, b in rows:
        e = tot.setdefault(k, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += v
        e[2] = max(e[2], v)
    s = sum(e[1] for e in tot.values())
    out = ['# %s: %d launches, %.1f us of device time (cold-cache, serialised: compare shares)' % (path, len(rows), s),
           '', '| kernel | launches | total us | max us | share |', '|---|---:|---:|---:|---:|']
    for k, e in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append('| %s | %d | %.1f | %.1f | %.1f%% |' % (k, e[0], e[1], e[2], 100 * e[1] / s))
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
