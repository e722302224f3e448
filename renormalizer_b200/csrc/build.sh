#!/bin/bash
# Build librn_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../librn_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="api.cu gemm_f64.cu pack.cu wapply.cu vecops.cu krylov.cu qr.cu qr_panel.cu svd_jacobi.cu hop.cu davidson.cu"
[ -f "$HERE/ozaki_gemm.cu" ] && SRCS="$SRCS ozaki_gemm.cu"
cd "$HERE"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -I"$HERE/../../include" -I"$HERE" \
  ${RN_NVCC_EXTRA} $SRCS -o "$OUT" -lcudart
echo "built $OUT"
