#!/bin/bash
# usage: build_variant.sh <out.so> <extra nvcc flags...>
set -e
OUT=$1; shift
cd /root/repo/probability_b200
TMP=$(mktemp -d)
for f in pb2_chain_kernels pb2_misc pb2_capi pb2_rowshard pb2_dense_tc pb2_tile pb2_tile_nuts pb2_logistic_tc pb2_comm; do
  [ -f csrc/$f.cu ] || continue
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$f.cu -o $TMP/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $TMP/*.o -ldl
rm -rf $TMP
echo built $OUT
