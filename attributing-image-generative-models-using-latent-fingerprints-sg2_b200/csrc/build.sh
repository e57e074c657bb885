#!/bin/bash
# Build liblfp_sg2.so for sm_100a (no torch dependency).  Usage: csrc/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lfp_native/liblfp_sg2.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS="runtime.cu bias_act.cu upfirdn2d.cu synth_kernels.cu synth_plan.cu attrib_kernels.cu"
SRCS="$SRCS conv_tc.cu modconv_plan.cu attrib_step.cu lpips_plan.cu mapping_pca.cu"
cd "$HERE"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared $SRCS -o "$OUT" -lcuda "$@"
echo "built $OUT"
