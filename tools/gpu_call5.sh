#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"house_panel|wy_update|wy_dots" -s 9 -c 3 -o gpurun_out/s6_qr python tools/qr_time.py 2048 256 > gpurun_out/s6_ncu.log 2>&1
tail -5 gpurun_out/s6_ncu.log
ls -la gpurun_out/
