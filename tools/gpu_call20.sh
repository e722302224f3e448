#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --profile-from-start off --cache-control none --clock-control none --import-source on -k regex:"ozaki_gemm_kernel" -s 2 -c 2 -o gpurun_out/s20_gemm python tools/site_update.py 256 > gpurun_out/s20_ncu.log 2>&1
tail -2 gpurun_out/s20_ncu.log
