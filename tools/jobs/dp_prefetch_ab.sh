#!/bin/bash
# N-GPU A/B of the host-batch staging in train_one_epoch (DDPM_B200_PREFETCH=0/1), short bench lines
N=${1:-2}
cd "$(dirname "$0")/../.."
for r in 1 2; do for p in ${ORDER:-0 1}; do
DDPM_B200_PREFETCH=$p python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$p bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --no-eager --no-c256 --no-cpu --no-ddim 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefetch $p: value', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'sync', round(d['e2e']['sync_every_step']['ms_per_step'],3))"
done; done
