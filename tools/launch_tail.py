#!/usr/bin/env python
"""Print the last N kernel launches (name, duration) of an `ncu --csv --metrics gpu__time_duration.sum` launch list."""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for r in rows[-n:]:
    print("  %-44s %8.2f us" % (r[ki].split("(")[0][-44:], float(r[vi].replace(",", "")) / 1000))
