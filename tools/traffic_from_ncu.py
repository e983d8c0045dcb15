"""Per kernel class DRAM traffic per launch from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum --csv` launch list of bench.py (the command is in profiles/r02_traffic.json["command"]).
bench.py reads the JSON this writes to fill roofline.traffic for the dominant class.

    python tools/traffic_from_ncu.py gpurun_out/r2_traffic.csv > profiles/r02_traffic.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict

CLASSES = [('conv_tcp_kernel', 'conv_tcp'), ('wgrad_finalize', 'wgrad_finalize'), ('wgrad', 'wgrad'), ('bn_bwd', 'bn_bwd'),
           ('bn_apply', 'bn_apply'), ('conv_ksplit', 'conv_train'), ('conv_patch', 'conv_eval'), ('conv_kernel', 'conv_train'),
           ('stem_kernel', 'conv_train'), ('knn_sv', 'knn_sv'), ('supcon', 'supcon'), ('pack_kernel', 'pack')]
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}


def main():
    with open(sys.argv[1], newline='') as fh:
        lines = [l for l in fh if not l.startswith('==')]
    per_launch = OrderedDict()
    for r in csv.DictReader(lines):
        key = r['ID']
        d = per_launch.setdefault(key, {'name': r['Kernel Name']})
        val = float(r['Metric Value'].replace(',', '')) * SCALE.get(r.get('Metric Unit', ''), 1.0)
        d[r['Metric Name']] = val
    agg = OrderedDict()
    for d in per_launch.values():
        cls = next((c for pat, c in CLASSES if pat in d['name']), None)
        if cls is None:
            continue
        a = agg.setdefault(cls, {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0, 'time_us': 0.0})
        a['launches'] += 1
        a['dram_read_bytes'] += d.get('dram__bytes_read.sum', 0.0)
        a['dram_write_bytes'] += d.get('dram__bytes_write.sum', 0.0)
        a['time_us'] += d.get('gpu__time_duration.sum', 0.0)
    out = {'command': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 '
                      '--csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline',
           'note': 'averages per launch over the first 1500 launches of the process (cold caches under ncu)', 'classes': {}}
    for cls, a in agg.items():
        n = a['launches']
        out['classes'][cls] = {'launches': n, 'dram_read_bytes_per_launch': a['dram_read_bytes'] / n,
                               'dram_write_bytes_per_launch': a['dram_write_bytes'] / n, 'time_us_per_launch': a['time_us'] / n}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
