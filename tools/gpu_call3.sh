#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/hop_roofline.py 1024 > gpurun_out/s4_hop1024.log 2>&1
cat gpurun_out/s4_hop1024.log
timeout 300 python tools/site_update.py 256 > gpurun_out/s4_site256.log 2>&1
cat gpurun_out/s4_site256.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s4_ncu.log 2>&1
tail -3 gpurun_out/s4_ncu.log; wc -l gpurun_out/s4_site256_launches.csv
