#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -k regex:"krylov_coef|gemm_tn_f64" --csv --log-file gpurun_out/s33_coef.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-roofline --no-cpu-baseline --profiler-range > gpurun_out/s33_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s33_coef.csv | head -8
for i in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-roofline | cut -c1-120; done
