"""Aggregate an `ncu --csv` launch list (gpu__time_duration + dram bytes) into one line per kernel."""
import collections
import csv
import io
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        name = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        unit, m = row['Metric Unit'], row['Metric Name']
        if m == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}[unit]
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        agg[name][m].append(v)
    print('%-34s %5s %10s %10s %10s %9s' % ('kernel', 'n', 'us/launch', 'rd MB', 'wr MB', 'GB/s'))
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
        t, rd, wr = d['gpu__time_duration.sum'], d.get('dram__bytes_read.sum', [0]), d.get('dram__bytes_write.sum', [0])
        n = len(t)
        print('%-34s %5d %10.1f %10.2f %10.2f %9.1f' % (k[:34], n, sum(t) / n, sum(rd) / n / 1e6, sum(wr) / n / 1e6,
                                                      (sum(rd) + sum(wr)) / sum(t) / 1e3))


if __name__ == '__main__':
    main(sys.argv[1])
