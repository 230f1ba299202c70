"""Print the headline fields of bench.py JSON lines (stdin or files)."""
import json, sys


def find(o, k):
    if isinstance(o, dict):
        if k in o:
            return o[k]
        for v in o.values():
            r = find(v, k)
            if r is not None:
                return r
    return None


for src in (sys.argv[1:] or ["-"]):
    txt = sys.stdin.read() if src == "-" else open(src).read()
    for line in txt.strip().splitlines():
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        rl = d.get("roofline") or {}
        print(src, "| gpus", d.get("n_gpus"), "| value %.4g" % d.get("value", 0), "| ms/step %.4f" % d.get("ms_per_step", 0), "| steps", d.get("steps"),
              "| e2e %.4g" % ((d.get("e2e") or {}).get("value") or 0), "| graph", find(d, "cuda_graph"), "| frl us", find(d, "round_us"), find(d, "transport"),
              "| roofline %.3f (%s)" % (rl.get("frac") or 0, rl.get("learn_ms")), "| clocks", d.get("clocks"))
