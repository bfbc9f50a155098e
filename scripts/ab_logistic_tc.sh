#!/bin/bash
# logistic tensor-core pipeline A/B on one box: gpurun_variants/libpb2_head.so (reference build) vs the in-tree library
for i in 1 2; do
  PB2_LIB_PATH=$PWD/gpurun_variants/libpb2_head.so timeout 200 python scripts/perf_rowshard.py 2>&1 | grep "tcgen05\|HMC" | sed 's/^/HEAD /'
  timeout 200 python scripts/perf_rowshard.py 2>&1 | grep "tcgen05\|HMC" | sed 's/^/NEW  /'
done
PB2_LIB_PATH=$PWD/gpurun_variants/libpb2_head.so timeout 200 python scripts/perf_logistic_tc.py 2>&1 | grep tcgen05 | sed 's/^/HEAD /'
timeout 200 python scripts/perf_logistic_tc.py 2>&1 | grep tcgen05 | sed 's/^/NEW  /'
