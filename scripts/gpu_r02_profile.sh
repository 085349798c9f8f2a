#!/bin/bash
# Round-2 evidence run (one GPU): ncu --set full captures of the attention kernels at 577 and 197 tokens (Nq = N launches),
# ncu launch lists (time + DRAM bytes per launch) of the ViT-L/384 and ViT-L/224 bench commands, per-kernel inference breakdowns.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
FULL="--set full --clock-control none --import-source on"
timeout 600 ncu $FULL -k regex:attn_.*_long_kernel -s 6 -c 4 -o $O/r02_attn_long -f python scripts/gpu_attn_prof.py 128 577 16 > $O/r02_ncu_attn_long.log 2>&1
timeout 600 ncu $FULL -k regex:attn_.*_fused_kernel -s 6 -c 4 -o $O/r02_attn_fused -f python scripts/gpu_attn_prof.py 256 197 16 > $O/r02_ncu_attn_fused.log 2>&1
LM="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
B="--steps 1 --warmup 3 --no-inference --no-cpu-baseline --no-torch-gpu"
timeout 900 ncu $LM -s 1460 -c 1000 --csv --log-file $O/r02_launches_vitl384_v1.csv python bench.py --workload vitl384 $B > $O/r02_ncu_l384.log 2>&1
timeout 900 ncu $LM -s 1400 -c 1000 --csv --log-file $O/r02_launches_vitl224_v1.csv python bench.py $B > $O/r02_ncu_l224.log 2>&1
for b in 1 8 64; do timeout 120 python scripts/gpu_infer_profile.py $b > $O/r02_infer_profile_b$b.log 2>&1; done
ls -la $O | tail -20
