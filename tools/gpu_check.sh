#!/bin/bash
# Quick GPU check (about 2 min): the parity suite and the host profile of a converged Holstein sweep.
( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 240 python tools/pyprof_dmrg.py 512 20 holstein_dmrg 3 > gpurun_out/r2f_dmrg_pyprof.txt 2>&1; head -22 gpurun_out/r2f_dmrg_pyprof.txt | cut -c1-220
