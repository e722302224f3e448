#!/bin/bash
# Routine GPU check (1 GPU): parity tests, smoke, one site update, bench line.
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/check_pytest.log 2>&1
tail -6 gpurun_out/check_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python tools/site_update.py 256
timeout 900 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
cat gpurun_out/check_bench.json; tail -2 gpurun_out/check_bench.err
