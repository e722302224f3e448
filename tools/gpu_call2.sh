#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_pytest.log 2>&1
tail -15 gpurun_out/s3_pytest.log
timeout 300 python tools/breakdown.py 1 256 > gpurun_out/s3_breakdown256.log 2>&1
cat gpurun_out/s3_breakdown256.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
cat gpurun_out/s3_bench.json; tail -3 gpurun_out/s3_bench.err
timeout 300 python tools/hop_roofline.py 1024 > gpurun_out/s3_hop1024.log 2>&1
cat gpurun_out/s3_hop1024.log
