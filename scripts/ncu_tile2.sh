#!/bin/bash

ncu --set full --clock-control none --import-source on -k regex:tile_nuts_kernel -s 0 -c 1 -o gpurun_out/prof_r01_tile_nuts_v2 python scripts/test_tile_nuts.py perf > gpurun_out/prof_tile_nuts_v2.log 2>&1
tail -n 3 gpurun_out/prof_tile_nuts_v2.log
