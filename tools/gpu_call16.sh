#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sweeps.py -m gpu -x -q ) > gpurun_out/s16_pytest.log 2>&1
tail -4 gpurun_out/s16_pytest.log
timeout 300 python tools/pyprof.py > gpurun_out/s16_pyprof.log 2>&1
head -45 gpurun_out/s16_pyprof.log
timeout 900 python bench.py --workload holstein_dmrg --bond 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s16_dmrg512.json 2> gpurun_out/s16_dmrg512.err
cat gpurun_out/s16_dmrg512.json; tail -3 gpurun_out/s16_dmrg512.err
