"""Instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K` (executed warp instructions and stall samples per opcode)."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
    hdr = rows[h]
    iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = 0
    byop, samp = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        n = int(r[iE])
        tot += n
        t = r[iS].strip().split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]
        byop[op] += n
        samp[op] += int(r[iSm] or 0)
    print('total executed warp instructions %d, static %d' % (tot, len(rows) - h - 1))
    for k, v in byop.most_common(top):  # noqa
        print('%-10s %12d %5.1f%%  stall samples %d' % (k, v, 100.0 * v / tot, samp[k]))


if __name__ == '__main__':
    main(sys.argv[1])
