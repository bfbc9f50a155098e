#!/bin/bash
# same-box A/B of library variants on C3 (logistic NUTS + DualAveraging)
for rep in 1 2; do for f in probability_b200/_C/ab_*.so; do
echo "$f: $(PB2_LIB_PATH=$PWD/$f timeout 200 python scripts/bench_configs.py c3 2>&1 | grep '^C3' | sed 's/.*chains: //')"
done; done
