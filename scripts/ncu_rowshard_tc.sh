#!/bin/bash
# full ncu capture of the tcgen05 row-sharded logistic gradient kernel (C5 shard size), 3rd launch of perf_rowshard.py
ncu --set full --clock-control none --import-source on -k regex:logistic_tc_kernel -s 2 -c 1 -o gpurun_out/r01_rowshard_tc python scripts/perf_rowshard.py > gpurun_out/r01_rowshard_tc.log 2>&1
tail -n 2 gpurun_out/r01_rowshard_tc.log
