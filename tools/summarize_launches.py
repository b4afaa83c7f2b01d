"""Aggregate an `ncu --csv` launch list (gpu__time_duration [+ dram bytes]) into one line per kernel, with each kernel's share of the
summed kernel time (under ncu launches are serialised and cold-cache: the SHARE is what compares with a timed run, not the absolute)."""
import collections
import csv
import io
import re
import sys


def kernel_name(full):
    name = re.sub(r'\((?:int|bool|unsigned int|long|unsigned long|char)\)', '', full)      # k_gemm<(int)1, (int)0> -> k_gemm<1, 0>
    name = name.split('(')[0].strip()
    return re.sub(r'^void\s+', '', name)


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        name = kernel_name(row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit, m = row['Metric Unit'], row['Metric Name']
        if m == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 's': 1e6, 'second': 1e6}[unit]
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]
        agg[name][m].append(v)
    total = sum(sum(d['gpu__time_duration.sum']) for d in agg.values())
    launches = sum(len(d['gpu__time_duration.sum']) for d in agg.values())
    print('%d launches, %.2f ms of kernel time in total' % (launches, total / 1e3))
    print('%-58s %6s %11s %10s %7s %10s %10s %9s' % ('kernel', 'n', 'us/launch', 'total ms', 'share', 'rd MB', 'wr MB', 'GB/s'))
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
        t, rd, wr = d['gpu__time_duration.sum'], d.get('dram__bytes_read.sum', [0]), d.get('dram__bytes_write.sum', [0])
        n = len(t)
        print('%-58s %6d %11.1f %10.2f %6.1f%% %10.2f %10.2f %9.1f' % (k[:58], n, sum(t) / n, sum(t) / 1e3, 100.0 * sum(t) / total, sum(rd) / n / 1e6,
                                                                      sum(wr) / n / 1e6, (sum(rd) + sum(wr)) / sum(t) / 1e3))


if __name__ == '__main__':
    main(sys.argv[1])
