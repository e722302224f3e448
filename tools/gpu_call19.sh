#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ozaki or hop" 2>&1 | tail -2
timeout 120 python tools/site_update.py 256
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s19_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s19_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s19_site256_launches.csv > gpurun_out/s19_summary.md; head -18 gpurun_out/s19_summary.md
timeout 600 ncu --set full --profile-from-start off --cache-control none --clock-control none --import-source on -k regex:"wapply_split|ozaki_split_t" -s 4 -c 2 -o gpurun_out/s19_small python tools/site_update.py 256 > gpurun_out/s19_ncu2.log 2>&1
