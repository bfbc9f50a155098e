#!/bin/bash
# full ncu capture (with source) of the async tile NUTS kernel (second fused launch of scripts/perf_tile_nuts.py)
PB2_ONLY_SCHED=1 ncu --set full --clock-control none --import-source on -k regex:tile_nuts_async_kernel -s 1 -c 1 -o gpurun_out/prof_async python scripts/perf_tile_nuts.py 16384 6 > gpurun_out/prof_async.log 2>&1
tail -n 2 gpurun_out/prof_async.log
