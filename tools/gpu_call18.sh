#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ozaki or hop" 2>&1 | tail -2
timeout 120 python tools/site_update.py 256
timeout 300 python tools/breakdown.py 1 256 2>&1 | head -3
timeout 300 python tools/hop_roofline.py 1024 2>&1 | tail -1
timeout 300 python tools/hop_roofline.py 512 2>&1 | tail -1
