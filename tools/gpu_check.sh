#!/bin/bash
# Quick GPU check (about 2 min): the parity suite, the smoke test and the headline bench line.
( time timeout 400 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --no-sub --steps 3 --warmup 3 > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; tail -2 gpurun_out/check_bench.err; python tools/show_bench.py gpurun_out/check_bench.json | cut -c1-300
