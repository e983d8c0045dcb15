"""Opcode histogram per kernel of the in-tree objects (cuobjdump -sass): the mnemonics that prove which hardware path
a kernel uses -- UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk, TMA 1-D),
UTMALDG / UTMASTG (cp.async.bulk.tensor), SYNCS (mbarrier), LDGSTS (cp.async), HMMA / IMMA (legacy mma.sync), FFMA.

    python tools/sass_histogram.py > profiles/r02_sass_histograms.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, 'online-continual-learning_b200', 'build')
WANT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'LDGSTS', 'HMMA', 'IMMA',
        'FFMA', 'DFMA', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'BAR', 'ATOM', 'RED']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r'\(.*', '', o.replace('b200ocl::(anonymous namespace)::', '').replace('void ', '')) for o in out]


def main():
    print('# SASS opcode histograms (cuobjdump -sass on online-continual-learning_b200/build/*.o, sm_100a)\n')
    print('Counts are static instruction counts per kernel; `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UBLKCP` = '
          'cp.async.bulk (TMA 1-D bulk copy), `UTMALDG` = cp.async.bulk.tensor (none: operands that need a layout change are '
          'staged by loader warps, see DESIGN.md section 5), `SYNCS` = mbarrier ops, `LDGSTS` = cp.async.\n')
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith('.o'):
            continue
        txt = subprocess.run(['cuobjdump', '-sass', os.path.join(OBJ, f)], capture_output=True, text=True).stdout
        kernels = re.split(r'\n\s*Function : ', txt)[1:]
        if not kernels:
            continue
        print('## %s\n' % f)
        print('| kernel | ' + ' | '.join(WANT) + ' | total |')
        print('|---|' + '---:|' * (len(WANT) + 1))
        names = demangle([k.split('\n', 1)[0].strip() for k in kernels])
        for name, k in zip(names, kernels):
            ops = re.findall(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', k, flags=re.M)
            hist = collections.Counter(o for o in ops)
            row = [sum(v for o, v in hist.items() if o.startswith(w)) for w in WANT]
            print('| `%s` | ' % name[:70] + ' | '.join(str(v) if v else '' for v in row) + ' | %d |' % len(ops))
        print()


if __name__ == '__main__':
    main()
