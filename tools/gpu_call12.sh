#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s12_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-roofline --no-cpu-baseline --profiler-range > gpurun_out/s12_ncu.log 2>&1
tail -2 gpurun_out/s12_ncu.log | cut -c1-300
python tools/launch_summary.py gpurun_out/s12_bench_launches.csv > gpurun_out/s12_summary.md
cat gpurun_out/s12_summary.md
