#!/bin/bash
# Sizes the cost of the gradient all-reduce at N GPUs: overlapped buckets vs one all-reduce at the end vs none (diagnostic).
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-inference --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1:], round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" "$@"; }
for rep in 1 2; do
run --bucket-mb 96
run --bucket-mb 100000
run --diag-no-allreduce
done
