#!/bin/bash
# round-2 evidence (run under gpurun, 1 GPU): the bench lines (never under a profiler), then
#  1. launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised)
#  2. ncu --set full of the dominant kernels: C2 asynchronous-lane tile NUTS (the fused warm-up launch of the bench), C4
#     stochastic-volatility NUTS (two CTAs per SM), C5 row-sharded logistic gradient (TMA-staged)
#  3. compute-sanitizer memcheck over the tensor-core / TMA / run-time-compiled kernels
set -x
O=gpurun_out/r02
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv \
  python bench.py --steps 20 --warmup 5 --no-ess --no-cpu-baseline --configs none > $O/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_nuts_async_kernel -s 156 -c 1 -o $O/tile_nuts_async \
  python bench.py --steps 20 --warmup 5 --no-ess --no-cpu-baseline --configs none > $O/tile_nuts_async.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 3 -c 2 -o $O/nuts_sv \
  python scripts/ncu_sv_nuts.py > $O/nuts_sv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logistic_tc_kernel -s 2 -c 1 -o $O/rowshard_tc \
  python scripts/perf_rowshard.py > $O/rowshard_tc.log 2>&1
compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $O/sanitizer_memcheck.log 2>&1
tail -n 4 $O/sanitizer_memcheck.log
tail -c 600 $O/bench_n1.err
