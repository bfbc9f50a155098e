#!/bin/bash
# C5 (rows sharded over N GPUs) with the cross-rank gradient sum as an NCCL all-reduce (0) vs fused into the step kernel
# over peer memory (1): usage scripts/ab_c5_collective.sh N
N=${1:-2}
for mode in 0 1 0 1; do
  PB2_ROWSHARD_COLLECTIVE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29600 + mode)) bench.py --gpus $N --steps 5 --warmup 3 --no-ess --no-cpu-baseline --configs c5 2>/dev/null | \
    python -c "
import sys, json
for line in sys.stdin:
  line = line.strip()
  if line.startswith('{'):
    d = json.loads(line)
    c = d['per_config']['c5']
    print('rowshard_collective=$mode N=$N: %.3f ms per transition, %.4g grad-evals/s, accept %.2f, shard_parity %s' % (
        c['ms_per_transition'], c['value'], c['accept_rate'], d.get('shard_parity')))
"
done
