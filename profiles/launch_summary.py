#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == 'ID':
            hdr = r
            continue
        if hdr is None:
            continue
        name = r[hdr.index('Kernel Name')]
        val = float(r[hdr.index('Metric Value')].replace(',', ''))
        unit = r[hdr.index('Metric Unit')]
        if unit == 'ns':
            val /= 1e3
        elif unit == 'ms':
            val *= 1e3
        agg[name][0] += 1
        agg[name][1] += val
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} us {v[0]:4d} {v[1] / v[0]:10.1f} us/launch {100 * v[1] / tot:5.1f}%  {k[:100]}")


if __name__ == '__main__':
    main()
