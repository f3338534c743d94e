#!/usr/bin/env python
"""One-line summary (+ per-kernel ms/step with -k) of bench.py JSON lines: python scripts/bsum.py [-k] file..."""
import json, sys
args = sys.argv[1:]
k = "-k" in args
for f in [a for a in args if a != "-k"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    g = d.get("roofline", {}).get("groups_ms_per_step", {})
    print(f, round(d["value"]), "seq/s", round(d["ms_per_step"], 3), "ms", "e2e", round(d.get("e2e", {}).get("value", 0)),
          "launches", d.get("gpu_launches"), "clk", d["clocks"]["sm_mhz"], {a: round(b, 2) for a, b in g.items()})
    if k:
        steps = d["steps"]
        for e in d.get("kernels", []):
            print("   %-34s %3d  %.3f ms/step" % (e["name"], e["launches"] // steps, e["ms"] / steps))
