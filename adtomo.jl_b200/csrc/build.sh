#!/bin/bash
# Builds libadtomo_b200.so for sm_100a (B200).  -fmad=false: the forward update must not be
# contracted into FMAs (bit parity with the reference's x86-64 build, see eik_core.h).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
      -Xcompiler -fPIC -shared ${NVCC_EXTRA} -o ../libadtomo_b200.so capi.cu -lcudart -ldl
