#!/bin/bash
# First GPU call of a session: parity tests, default bench, per-piece breakdown, H_eff roofline.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2_pytest.log 2>&1
tail -5 gpurun_out/s2_pytest.log
timeout 600 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
cat gpurun_out/s2_bench.json
timeout 300 python tools/breakdown.py 1 256 > gpurun_out/s2_breakdown256.log 2>&1
cat gpurun_out/s2_breakdown256.log
timeout 300 python tools/hop_roofline.py 1024 > gpurun_out/s2_hop1024.log 2>&1
cat gpurun_out/s2_hop1024.log
timeout 300 python tools/hop_roofline.py 512 > gpurun_out/s2_hop512.log 2>&1
cat gpurun_out/s2_hop512.log
timeout 300 python tools/pyprof.py > gpurun_out/s2_pyprof.log 2>&1
head -50 gpurun_out/s2_pyprof.log
