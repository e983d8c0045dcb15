"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python tools/summarize_launches.py gpurun_out/launches_r01.csv [skip_first_n] > profiles/...md"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline='') as fh:
        lines = [l for l in fh if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = r['Kernel Name']
        val = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
        rows.append((name, val * scale, r.get('Grid Size', ''), r.get('Block Size', '')))
    agg = OrderedDict()
    for name, us, grid, block in rows:
        short = re.sub(r'\(.*$', '', name)
        short = re.sub(r'^void ', '', short)
        short = re.sub(r'b200ocl::\(anonymous namespace\)::', '', short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f%% | %.2f |' % (k[:90], n, us, 100 * us / total, us / n))
    print('| **total** | %d | %.1f | 100%% | |' % (sum(v[0] for v in agg.values()), total))


if __name__ == '__main__':
    main()
