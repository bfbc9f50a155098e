#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:tile_hmc_kernel -s 2 -c 1 -o gpurun_out/prof_r01_tile_hmc_v2 python scripts/test_tile_hmc.py > gpurun_out/prof_tile_hmc_v2.log 2>&1
tail -n 2 gpurun_out/prof_tile_hmc_v2.log
