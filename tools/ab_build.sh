#!/bin/bash
# Build A/B variants of the library that differ in -D flags of the GEMM translation unit:
#   tools/ab_build.sh name "-DFLAG=1 ..."   ->  protoquant_b200/libpq_<name>.so   (select with PQ_LIB_PATH)
set -e
cd "$(dirname "$0")/../protoquant_b200/csrc"
name=$1; flags=$2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $flags -c qgemm_tcgen05.cu -o /tmp/qgemm_$name.o 2>/tmp/qgemm_$name.log
$NVCC $ARCH -shared -o ../libpq_$name.so capi.o rowwise_quant.o dequant.o fused_quant.o rowparallel.o /tmp/qgemm_$name.o qgemm_smallm.o -cudart shared
echo built ../libpq_$name.so
