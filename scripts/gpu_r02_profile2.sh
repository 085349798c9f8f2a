#!/bin/bash
# Round-2 final evidence (one GPU): launch lists (time + DRAM bytes per launch) of the ViT-L/224 and ViT-L/384 bench commands with
# the final kernels, ncu --set full of the 577-token attention kernels and of the residual / GELU GEMMs, ViT-B bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
FULL="--set full --clock-control none --import-source on"
LM="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
B="--steps 1 --warmup 3 --no-inference --no-cpu-baseline --no-torch-gpu"
timeout 900 ncu $LM -s 1400 -c 1000 --csv --log-file $O/r02_launches_vitl224_v2.csv python bench.py $B > $O/r02_ncu_l224b.log 2>&1
timeout 900 ncu $LM -s 1460 -c 1000 --csv --log-file $O/r02_launches_vitl384_v2.csv python bench.py --workload vitl384 $B > $O/r02_ncu_l384b.log 2>&1
timeout 600 ncu $FULL -k regex:attn_.*_long_kernel -s 6 -c 2 -o $O/r02_attn_long_v2 -f python scripts/gpu_attn_prof.py 128 577 16 > $O/r02_ncu_attn_long2.log 2>&1
timeout 600 ncu $FULL -k regex:attn_.*_fused_kernel -s 6 -c 2 -o $O/r02_attn_fused_v6 -f python scripts/gpu_attn_prof.py 256 197 16 > $O/r02_ncu_attn_fused6.log 2>&1
timeout 600 ncu $FULL -k regex:gemm_bf16_tcgen05_kernel -s 3 -c 48 -o $O/r02_gemm_epi_v2 -f python scripts/gpu_gemm_epi_prof.py > $O/r02_ncu_gemm_epi2.log 2>&1
timeout 300 python bench.py --workload vitb224 --no-cpu-baseline --no-torch-gpu > $O/r02_bench_vitb224.json 2> $O/r02_bench_vitb224.err
ls -la $O | grep -E "v2|v6|vitb" 
