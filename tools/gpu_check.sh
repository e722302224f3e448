#!/bin/bash
# Quick GPU check: the parity suite, then the ab initio DMRG workload as the headline of a short bench run.
( time timeout 400 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 200 python tools/hop_qc_shape.py 2>&1 | tail -1
timeout 400 python bench.py --workload qc_dmrg --steps 1 --warmup 1 --no-e2e --no-roofline > gpurun_out/r2f_bench_qc.json 2> gpurun_out/r2f_bench_qc.err; tail -2 gpurun_out/r2f_bench_qc.err; python tools/show_bench.py gpurun_out/r2f_bench_qc.json | cut -c1-400
