#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_sweeps.py::test_dmrg_omega_targeting_golden ) > gpurun_out/s24_pytest.log 2>&1
tail -6 gpurun_out/s24_pytest.log
timeout 120 python tools/site_update.py 256
RN_KRYLOV_FUSED_TAIL=0 timeout 120 python tools/site_update.py 256
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s24_bench.json 2> gpurun_out/s24_bench.err
cat gpurun_out/s24_bench.json; tail -3 gpurun_out/s24_bench.err
RN_KRYLOV_FUSED_TAIL=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-roofline | cut -c1-200
