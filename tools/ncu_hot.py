"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` output.
    ncu -i x.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > /tmp/src.csv
    python tools/ncu_hot.py /tmp/src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
def _num(x):
    try:
        float(x or 0)
        return True
    except ValueError:
        return False
body = [r for r in rows[2:] if len(r) == len(hdr) and _num(r[isamp]) and r[ia].startswith("0x") is not None and _num(r[iex])]
tot = sum(float(r[isamp] or 0) for r in body)
print("kernel:", rows[0][1][:100], " total samples:", tot)
order = sorted(range(len(body)), key=lambda i: -float(body[i][isamp] or 0))[:N]
for i in sorted(order):
    r = body[i]
    st = sorted(((float(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% ex=%9s  %-70s %s" % (i, 100 * float(r[isamp] or 0) / tot, r[iex], r[isrc][:70], " ".join("%s=%d" % (n, v) for v, n in st if v)))
