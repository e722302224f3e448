#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/site_update.py 256
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s17_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s17_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s17_site256_launches.csv 0 30 > gpurun_out/s17_summary.md; cat gpurun_out/s17_summary.md
