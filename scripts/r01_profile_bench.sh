#!/bin/bash
# round-1 evidence for bench.py's dominant kernel (run under gpurun, 1 GPU):
#  1. launch list of the bench command (gpu__time_duration per kernel launch; cold-cache, serialised)
#  2. one --set full capture of the timed tile_nuts_async_kernel launch (index 154 = 150 adaptation steps + 3 warm-up
#     one_step launches + the fused warm-up call)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01_launches_async.csv \
  python bench.py --steps 20 --warmup 3 --no-ess --no-cpu-baseline > gpurun_out/r01_launches_async.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_nuts_async_kernel -s 154 -c 1 -o gpurun_out/r01_tile_nuts_async \
  python bench.py --steps 20 --warmup 3 --no-ess --no-cpu-baseline > gpurun_out/r01_tile_nuts_async.log 2>&1
tail -n 3 gpurun_out/r01_tile_nuts_async.log | cut -c1-300
