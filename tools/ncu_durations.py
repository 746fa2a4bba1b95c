#!/usr/bin/env python
"""Prints the per-launch durations (us) of the kernels matching a regex from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`), in launch
order -- the way to time one kernel variant against another without host-side noise.
usage: tools/ncu_durations.py launches.csv REGEX"""
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
pat = re.compile(sys.argv[2])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum" or not pat.search(r["Kernel Name"]):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
    print("%9.2f  %s  grid=%s" % (v, re.sub(r"\(.*", "", r["Kernel Name"])[-60:], r.get("Grid Size", "")))
