#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and
mean device time and share of one evaluation.  Usage: summarize_launches.py launches.csv [evaluation index]"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"((?:[\w:]+::)?[\w]+)(<.*>)?\(", name)
    base = name.split("(")[0]
    base = base.replace("mpid::", "")
    return base[:110]


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    which = int(sys.argv[2]) if len(sys.argv) > 2 else None
    if which is not None:
        # evaluation `which` (0-based): from its first kernel -- k_wrap_cells when the evaluation sorts and searches,
        # k_regather_sites when it reuses the order and the candidate list -- to the first kernel of the next one
        starts = [i for i, r in enumerate(rows) if "k_wrap_cells" in r["Kernel Name"] or "k_regather_sites" in r["Kernel Name"]] + [len(rows)]
        rows = [r for r in rows[starts[which]:starts[which + 1]] if "at::native" not in r["Kernel Name"]]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | mean us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |" % (k, a[0], a[1]/1e3, a[1]/1e3/a[0], 100*a[1]/total, a[2], a[3]))
    print("\nlaunches: %d, summed device time: %.1f us (serialised, cold cache under ncu)" % (len(rows), total/1e3))


if __name__ == "__main__":
    main()
