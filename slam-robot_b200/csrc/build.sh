#!/bin/sh
# Builds libslamfe.so in-tree for sm_100a.  -fmad=false: no implicit FMA contraction (every
# fused multiply-add in the sources is an explicit fmaf), see sfe_common.cuh.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
  --shared -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v \
  -o libslamfe.so capi.cu replay.cu pyramid.cu pyramid_stream.cu track_hessian.cu klt.cu brute.cu hamming.cu hamming_mma.cu gftt.cu seed.cu dist.cu -ldl "$@"
