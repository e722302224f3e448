#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "qr" ) > gpurun_out/s9_pytest_qr.log 2>&1
tail -15 gpurun_out/s9_pytest_qr.log
timeout 120 python tools/site_update.py 256
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s9_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s9_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s9_site256_launches.csv | head -12
