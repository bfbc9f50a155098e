#!/bin/bash
# C2 kernel A/B on one box: gpurun_variants/libpb2_head.so (reference build) vs the in-tree library
export PB2_ONLY_SCHED=1
for i in 1 2; do
  PB2_LIB_PATH=$PWD/gpurun_variants/libpb2_head.so timeout 120 python scripts/perf_tile_nuts.py 16384 20 10 2>&1 | grep "variant 0" | sed 's/^/HEAD /'
  timeout 120 python scripts/perf_tile_nuts.py 16384 20 10 2>&1 | grep "variant 0" | sed 's/^/NEW  /'
done
