#!/bin/bash
# full ncu capture of the tcgen05 logistic gradient primitive (C3 shape, 65,536 chains): a late launch of perf_logistic_tc.py
ncu --set full --clock-control none --import-source on -k regex:logistic_tc_kernel -s 30 -c 1 -o gpurun_out/r01_logistic_tc python scripts/perf_logistic_tc.py > gpurun_out/r01_logistic_tc.log 2>&1
tail -n 2 gpurun_out/r01_logistic_tc.log
