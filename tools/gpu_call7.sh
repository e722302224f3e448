#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:"house_panel" -s 3 -c 1 -o gpurun_out/s10_panel python tools/qr_time.py 2048 256 > gpurun_out/s10_ncu.log 2>&1
tail -3 gpurun_out/s10_ncu.log
