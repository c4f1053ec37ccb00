"""Print the essentials of a bench.py JSON line read from stdin (helper for the gpu_*.sh scripts)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print("%s value %.4g  ms/step %.4f  kernel %s %.4f ms  %.1f GB/s  frac %.3f  e2e %.4g" % (
    tag, d["value"], d["ms_per_step"], r["kernel"], r["kernel_ms"], r["achieved"], r["frac"], d["e2e"]["value"]))
