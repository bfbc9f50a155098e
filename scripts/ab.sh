#!/bin/bash
# same-box A/B of prebuilt library variants (probability_b200/_C/ab_*.so)
for rep in 1 2; do for f in probability_b200/_C/ab_*.so; do for K in ${AB_KS:-20 100}; do
echo "$f K=$K: $(PB2_LIB_PATH=$PWD/$f PB2_ONLY_SCHED=1 timeout 100 python scripts/perf_tile_nuts.py 16384 $K 2>&1 | tail -1 | sed 's/.*best//')"
done; done; done
