#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --profile-from-start off --cache-control none --clock-control none --import-source on -k regex:"ozaki_gemm_kernel" -s 14 -c 4 -o gpurun_out/s13_gemm python tools/site_update.py 256 > gpurun_out/s13_ncu.log 2>&1
tail -3 gpurun_out/s13_ncu.log
