"""Prints the per-kernel table of a bench.py JSON line (ms per step, share)."""
import json, sys
d = json.load(open(sys.argv[1]))
steps = d["steps"]
tot = sum(k["ms"] for k in d["kernels"])
for k in sorted(d["kernels"], key=lambda k: -k["ms"]):
    print("%-32s %4d %8.3f ms/step %5.1f%%  %8.1f (GB/s | TFLOP/s)" % (k["name"], k["launches"] / steps, k["ms"] / steps, 100 * k["ms"] / tot,
          k["work"] / (k["ms"] * 1e-3) / (1e12 if k["name"].startswith("conv_") else 1e9)))
print("tagged total %.3f ms/step; step %.3f ms; value %.1f; e2e %.1f; launches %d" % (tot / steps, d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
print(d.get("clocks"))
