#!/bin/bash
# GPU job (N GPUs): data-parallel timeline + bench, graded gradient buckets vs round-1 fixed 16 MB buckets
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
$RUN tools/dp_timeline.py 128 6 > gpurun_out/r2_timeline_dp${N}.txt 2>gpurun_out/r2_timeline_dp${N}.err
DDPM_B200_DP_TAIL_MB=0 $RUN tools/dp_timeline.py 128 6 > gpurun_out/r2_timeline_dp${N}_fixed16.txt 2>>gpurun_out/r2_timeline_dp${N}.err
head -4 gpurun_out/r2_timeline_dp${N}.txt; grep -E "exposed|NCCL kernels" gpurun_out/r2_timeline_dp${N}.txt
head -3 gpurun_out/r2_timeline_dp${N}_fixed16.txt; grep -E "exposed|NCCL kernels" gpurun_out/r2_timeline_dp${N}_fixed16.txt
for tail in 12 0; do
  DDPM_B200_DP_TAIL_MB=$tail $RUN bench.py --gpus $N --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu > gpurun_out/r2_bench_dp${N}_tail$tail.json 2> gpurun_out/r2_bench_dp${N}_tail$tail.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_dp${N}_tail$tail.json").read().strip().splitlines()[-1])
    print("N=$N tail=$tail", d["value"], d["ms_per_step"], d["e2e"]["value"], d["ddim100"]["value"])
except Exception as e:
    print("N=$N tail=$tail failed", e)
PY
done
python tools/dp_timeline.py 128 6 2>/dev/null | head -3
