#!/bin/bash
# A/B on N GPUs: NCCL's CTA budget for the overlapped gradient all-reduce (fewer NCCL CTAs = fewer SMs taken from the backward)
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { tag=$1; shift; env "$@" timeout 300 $TR --master-port 29521 bench.py --gpus $N --steps 12 --warmup 3 --no-inference --no-cpu-baseline > $O/r02_nccl_$tag.json 2> $O/r02_nccl_$tag.err; python -c "
import json,sys; d=json.loads(open('$O/r02_nccl_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz e2e', round(d['e2e']['value'],1))"; }
run default A=1
run max8 NCCL_MAX_CTAS=8
run max4 NCCL_MAX_CTAS=4
run max2 NCCL_MAX_CTAS=2
run default2 A=1
