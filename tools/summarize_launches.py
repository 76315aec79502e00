#!/usr/bin/env python3
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`): per kernel total time, share,
launch count and average; kernel names shortened to their template head.  Usage: summarize_launches.py X.csv "title" """
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    ik, iv, im = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^(void )?(spirk::)?", "", r[ik])
        name = re.sub(r"\(.*$", "", name)
        tot[name] += float(r[iv].replace(",", "")) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    print(title)
    print(f"total {total / 1e3:.2f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare SHARES)\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{tot[k]:10.1f} us {100 * tot[k] / total:5.1f}% n={cnt[k]:5d} avg {tot[k] / cnt[k]:8.1f}  {k}")


if __name__ == "__main__":
    main()
