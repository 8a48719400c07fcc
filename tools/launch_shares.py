#!/usr/bin/env python
"""Kernel shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_shares.py profiles/r1_launches_v10_batch592.csv [skip_first_n_launches]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[skip:]:
        name, val, unit = r[4], float(r[-1].replace(",", "")), r[-2]
        ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
        tot[name.split("(")[0]] += ms
        cnt[name.split("(")[0]] += 1
    s = sum(tot.values())
    for k, v in tot.most_common():
        print(f"{v:10.3f} ms  {100 * v / s:5.1f} %  {cnt[k]:4d} launches  {k}")


if __name__ == "__main__":
    main()
