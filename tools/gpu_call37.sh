#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ozaki or hop" 2>&1 | tail -8
timeout 120 python tools/hop_roofline.py 1024 2>&1 | tail -2
RN_OZ_CTA2=0 timeout 120 python tools/hop_roofline.py 1024 2>&1 | tail -1
timeout 120 python tools/hop_roofline.py 256 2>&1 | tail -1
RN_OZ_CTA2=0 timeout 120 python tools/hop_roofline.py 256 2>&1 | tail -1
