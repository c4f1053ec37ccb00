"""Print the essentials of a bench.py JSON line read from stdin (helper for the gpu_*.sh scripts)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
except Exception as e:  # noqa: BLE001
    print(tag, "no bench line:", e)
    sys.exit(0)
r = d["roofline"]
print("%s value %.4g  ms/step %.4f  kernel %s %.4f ms  %.1f GB/s  frac %.3f  collective_ms %.4f  launch %s" % (
    tag, d["value"], d["ms_per_step"], r["kernel"], r["kernel_ms"], r["achieved"], r["frac"], d.get("collective_ms", -1), r["kernel_launch"].get("tile_order")))
e = d.get("e2e")
if e:
    print("   e2e %.4g (%s)" % (e["value"], e.get("which")))
    for k, v in (e.get("variants") or {}).items():
        if v:
            print("      %-28s %.4g  (%d steps, d2h %s B/step)" % (k, v["value"], v["steps"], v["d2h_bytes_per_step"]))
if d.get("config4"):
    c = d["config4"]
    print("   config4 value %.4g  frac %.3f  collective_ms %.4f" % (c["value"], c["roofline"]["frac"], c["collective_ms"]))
if d.get("config5"):
    c = d["config5"]
    print("   config5 value %.4g  frac %.3f  device_rollout %.4g" % (c["value"], c["roofline"]["frac"], c["device_rollout_mlp_policies"]["value"]))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    print("   cpu_baseline %.4g on %d cores (%s); python_reference %s" % (c["value"], c["cores"], c["kind"], (c.get("python_reference") or {}).get("value")))
print("   clocks", d.get("clocks"), "numa", d["config"].get("numa"))
