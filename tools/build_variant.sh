#!/bin/bash
# usage: tools/build_variant.sh TILE MINB [extra nvcc flags]  -> variants/lib_TILE_MINB[_tag].so   (dev aid)
T=$1; B=$2; shift 2
TAG=${TAG:-}
mkdir -p variants
cd microclimf_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include \
  -DMCF_TILE=$T -DMCF_MINB=$B "$@" -shared -o ../../variants/lib_${T}_${B}${TAG}.so mcf_kernels.cu mcf_terrain.cu mcf_snow.cu mcf_api.cu -Xptxas -v 2>&1 \
  | grep -A2 "k_gridILi0ELi0ELb0" | grep "Used\|spill" | tr '\n' ' '
echo " <= $T x $B $TAG"
