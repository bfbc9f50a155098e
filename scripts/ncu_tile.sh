#!/bin/bash
# one full ncu capture each of tile_hmc_kernel and tile_nuts_kernel
ncu --set full --clock-control none --import-source on -k regex:tile_hmc_kernel -s 1 -c 1 -o gpurun_out/prof_r01_tile_hmc python scripts/test_tile_hmc.py > gpurun_out/prof_tile_hmc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_nuts_kernel -s 1 -c 1 -o gpurun_out/prof_r01_tile_nuts python scripts/test_tile_nuts.py perf > gpurun_out/prof_tile_nuts.log 2>&1
tail -2 gpurun_out/prof_tile_hmc.log gpurun_out/prof_tile_nuts.log
