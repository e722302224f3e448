#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sweeps.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/site_update.py 256
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s34_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s34_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s34_site256_launches.csv > gpurun_out/s34_summary.md; head -16 gpurun_out/s34_summary.md
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-roofline | cut -c1-120; done
