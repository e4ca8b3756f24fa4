"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (last `frac` of the run)."""
import collections
import csv
import sys


def main(path, parts=3):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        seq.append((r[ki].split('(')[0][:70], v))
    n = len(seq) // parts
    last = seq[-n:]
    tot = sum(v for _, v in last)
    agg = collections.OrderedDict()
    for name, v in last:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.1f} us  {100 * v / tot:5.1f}%  x{c:3d}  {name}")
    print(f"total {tot:.1f} us over {len(last)} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
