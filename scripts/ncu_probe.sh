#!/bin/bash
# usage: scripts/ncu_probe.sh <cfg c1|c2|c3|c4> <out-name> [skip] ; one --set full capture of the NUTS chain kernel
cfg=$1; name=$2; skip=${3:-2}
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s $skip -c 1 -o gpurun_out/$name env PROBE_NOADAPT=1 python scripts/probe_perf.py $cfg > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
