"""Print the per-kernel durations of the LAST learn step recorded in an ncu launch list (csv from
`ncu --metrics gpu__time_duration.sum --csv`).   python tools/launches.py gpurun_out/x.csv [iters_in_file]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
rows = rows[1:]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = len(rows) // iters
tot = 0.0
for r in rows[(iters - 1) * n:]:
    us = float(r[vi].replace(",", "")) / (1000 if "ns" in r[ui] else 1)
    tot += us
    print("%8.1f us  %s" % (us, r[ki][:100]))
print("%8.1f us  TOTAL (%d launches)" % (tot, n))
