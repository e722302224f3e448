python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/int8_peak.py profiles/r02_int8_peak.json
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 2500 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench2.json"))
def show(r):
    print(r.get("workload", r.get("config",{}).get("workload")))
    print("  value", r["value"], "ms/step", r["ms_per_step"], "e2e", r["e2e"] and r["e2e"]["value"], "launches", r["gpu_launches"])
    print("  roofline", r.get("roofline") and {k:r["roofline"][k] for k in ("achieved","frac","launches","avg_launch_us","peak")})
    print("  cpu", r.get("cpu_baseline"))
    print("  parity", r.get("parity_check"))
show(d)
for s in d["sub_results"]: show(s)
PY
cp profiles/r02_int8_peak.json gpurun_out/
