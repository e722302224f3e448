"""Print the interesting fields of a bench.py JSON line (headline + sub_results + strong_scaling)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(r):
    print(r.get("workload", r.get("config", {}).get("workload")))
    if "error" in r:
        print("  ERROR", r["error"])
        return
    print("  value", round(r["value"], 3), "ms/step", round(r["ms_per_step"], 1), "e2e", r["e2e"] and round(r["e2e"]["value"], 3),
          "launches", r["gpu_launches"])
    rf = r.get("roofline")
    print("  roofline", rf and {k: rf[k] for k in ("achieved", "frac", "launches", "avg_launch_us", "peak")})
    c = r.get("cpu_baseline")
    print("  cpu", c and (round(c["value"], 4), c["sample"][:160]))
    p = r.get("parity_check")
    print("  parity", p and {k: v for k, v in p.items() if k not in ("what",)})
show(d)
for s in d.get("sub_results", []):
    show(s)
for s in d.get("strong_scaling", []):
    print({k: v for k, v in s.items() if k not in ("sharding", "collective", "energies")})
print("clocks", d.get("clocks"))
