#!/bin/bash
# scheduler policy sweep: grid size x patience (fused K=20 NUTS on the C2 workload)
for g in 148 132 120; do for pt in 8 24 96; do
echo "grid $g patience $pt: $(PB2_SCHED_GRID=$g PB2_SCHED_PATIENCE=$pt PB2_ONLY_SCHED=1 timeout 120 python scripts/test_tile_sched.py 16384 20 2>&1 | grep '^variant 0:')"
done; done
