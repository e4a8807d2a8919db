#!/bin/bash
# GPU job (8 GPUs): where the +0.5 ms of the 8-GPU step sits, and what moves it
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
DDPM_B200_DP_TAIL_MB=0 $RUN tools/dp_timeline.py 128 6 > gpurun_out/r2_timeline_dp${N}_fixed16.txt 2>gpurun_out/r2_timeline_dp${N}.err
$RUN tools/dp_timeline.py 128 6 > gpurun_out/r2_timeline_dp${N}.txt 2>>gpurun_out/r2_timeline_dp${N}.err
grep -E "world|buckets|exposed|NCCL kernels" gpurun_out/r2_timeline_dp${N}_fixed16.txt gpurun_out/r2_timeline_dp${N}.txt
run_bench() {  # name, env...
  name=$1; shift
  env "$@" $RUN bench.py --gpus $N --steps 20 --warmup 5 --no-c256 --no-eager --no-cpu --no-ddim > gpurun_out/r2_bench_dp${N}_$name.json 2> gpurun_out/r2_bench_dp${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_dp${N}_$name.json").read().strip().splitlines()[-1])
    print("N=$N $name", round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]))
except Exception as e:
    print("N=$N $name failed", e)
PY
}
run_bench fixed16 DDPM_B200_DP_TAIL_MB=0
run_bench graded DDPM_B200_DP_TAIL_MB=12
run_bench fixed16_maxctas4 DDPM_B200_DP_TAIL_MB=0 NCCL_MAX_CTAS=4
run_bench fixed16_maxctas8 DDPM_B200_DP_TAIL_MB=0 NCCL_MAX_CTAS=8
run_bench graded_maxctas4 DDPM_B200_DP_TAIL_MB=12 NCCL_MAX_CTAS=4
run_bench fixed16_nvls DDPM_B200_DP_TAIL_MB=0 NCCL_ALGO=NVLS
run_bench fixed16_tree DDPM_B200_DP_TAIL_MB=0 NCCL_ALGO=Tree
run_bench fixed16_b EMPTY=1 DDPM_B200_DP_TAIL_MB=0
