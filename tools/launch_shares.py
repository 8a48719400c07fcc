#!/usr/bin/env python
"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_shares.py CSV [first_run last_run]
A "run" is one solver run (it starts with a k_init launch).  With the default bench command the runs are: 0 the untimed
preparation (20 iterations), 1-3 warm-up ticks, 4-8 the timed device-resident ticks, 9-14 the end-to-end ticks, 15-20 the
closed-loop ticks; `... 4 9` therefore prints the per-kernel time of the timed region."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    if len(sys.argv) > 3:
        starts = [i for i, r in enumerate(rows) if r[4].startswith("k_init")] + [len(rows)]
        a, b = int(sys.argv[2]), int(sys.argv[3])
        rows, nruns = rows[starts[a]:starts[b]], b - a
    else:
        nruns = 1
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name, val, unit = r[4].split("(")[0], float(r[-1].replace(",", "")), r[-2]
        ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
        tot[name] += ms
        cnt[name] += 1
    s = sum(tot.values())
    for k, v in tot.most_common():
        print(f"{v / nruns:10.3f} ms/run  {100 * v / s:5.1f} %  {cnt[k] / nruns:6.1f} launches/run  {k}")
    print(f"{s / nruns:10.3f} ms/run  total")


if __name__ == "__main__":
    main()
