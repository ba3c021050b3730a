#!/usr/bin/env python
"""Summarise a MPIDB200_TRACE timeline: busy time per stream, idle gaps on the main stream, the longest gaps."""
import csv
import sys

rows = []
for line in list(open(sys.argv[1]))[1:]:
    # kernel names may contain commas (template arguments): split from both ends
    sid, rest = line.rstrip("\n").split(",", 1)
    name, a, b = rest.rsplit(",", 2)
    rows.append(dict(stream=sid, name=name, start_us=float(a), end_us=float(b)))
end = max(r["end_us"] for r in rows)
print("evaluation span %.1f us, %d launches" % (end, len(rows)))
for sid in ("1", "2", "3"):
    rs = sorted((r for r in rows if r["stream"] == sid), key=lambda r: r["start_us"])
    busy = sum(r["end_us"] - r["start_us"] for r in rs)
    print("stream %s: %d launches, busy %.1f us" % (sid, len(rs), busy))
# union of busy intervals over all streams -> time with nothing running
iv = sorted((r["start_us"], r["end_us"]) for r in rows)
cur_s, cur_e = iv[0]
idle = []
for a, b in iv[1:]:
    if a > cur_e:
        idle.append((cur_e, a))
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
print("GPU idle (no kernel on any stream): %.1f us in %d gaps" % (sum(b - a for a, b in idle), len(idle)))
print("largest gaps:")
for a, b in sorted(idle, key=lambda g: g[0] - g[1])[:12]:
    before = max((r for r in rows if r["end_us"] <= a + 1e-6), key=lambda r: r["end_us"])
    after = min((r for r in rows if r["start_us"] >= b - 1e-6), key=lambda r: r["start_us"])
    print("  %.1f us at t=%.1f  after %s  before %s" % (b - a, a, before["name"][:40], after["name"][:40]))
if len(sys.argv) > 2:
    for r in sorted(rows, key=lambda r: r["start_us"]):
        print("%s %8.1f %8.1f %6.1f  %s" % (r["stream"], r["start_us"], r["end_us"], r["end_us"] - r["start_us"], r["name"][:60]))
