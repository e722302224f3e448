#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "qr" ) > gpurun_out/s5_pytest_qr.log 2>&1
tail -15 gpurun_out/s5_pytest_qr.log
timeout 120 python tools/site_update.py 256 > gpurun_out/s5_site256.log 2>&1
cat gpurun_out/s5_site256.log
RN_QR_PANEL=0 timeout 120 python tools/site_update.py 256
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s5_site256_launches.csv python tools/site_update.py 256 > gpurun_out/s5_ncu.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s5_pytest.log 2>&1
tail -5 gpurun_out/s5_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
cat gpurun_out/s5_bench.json; tail -3 gpurun_out/s5_bench.err
