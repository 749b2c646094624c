"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv [n_steps]"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    m = re.match(r'(?:void\s+)?([A-Za-z_0-9:]+(?:<[0-9, ]+>)?)', name)      # keep small integer template arguments (kernel variants)
    base = m.group(1) if m else name
    if base.startswith('at::') or base.startswith('at_cuda'):
        # keep the functor for torch elementwise kernels
        f = re.search(r'(\w+(?:Functor|_kernel_cuda|kernel_impl|Op|Ops)\w*)', name[len(base):])
        base = base + ('<' + f.group(1) + '>' if f else '')
    return base[:90]


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
        k = short(row['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%d kernels, %.1f us total (%.1f us / step over %g steps)' % (sum(v[0] for v in agg.values()), tot, tot / steps, steps))
    print('%-90s %6s %11s %6s' % ('kernel', 'count', 'us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-90s %6d %11.1f %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == '__main__':
    main()
