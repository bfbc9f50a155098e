#!/bin/bash
# full ncu capture (with source) of the lock-step tile NUTS kernel: second launch of scripts/tile_prof.py
ncu --set full --clock-control none --import-source on -k regex:tile_nuts_kernel -s 1 -c 1 -o gpurun_out/prof_tile_lock python scripts/tile_prof.py > gpurun_out/prof_tile_lock.log 2>&1
tail -n 2 gpurun_out/prof_tile_lock.log
