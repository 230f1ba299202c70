"""Per-launch DRAM traffic and duration from an .ncu-rep (ncu --set full), as JSON for profiles/ and bench.py.
    python tools/ncu_traffic.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/r01_traffic.json"""
import csv, io, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    out = []
    for r in rows[2:]:
        val = lambda k: float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]]
        out.append({"kernel": r[col["Kernel Name"]].split("(")[0], "us_under_ncu": round(val("gpu__time_duration.sum"), 2),
                    "dram_read_bytes": int(val("dram__bytes_read.sum")), "dram_write_bytes": int(val("dram__bytes_write.sum"))})
    return out


if __name__ == "__main__":
    res = {}
    for rep in sys.argv[1:]:
        res[rep.split("/")[-1]] = launches(rep)
    json.dump(res, sys.stdout, indent=1)
