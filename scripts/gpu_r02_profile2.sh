#!/bin/bash
# Round-2 final evidence (one GPU): launch lists (time + DRAM bytes per launch) of the ViT-L/224 and ViT-L/384 bench commands with
# the final kernels, ncu --set full of the attention kernels and of one launch per GEMM variant, the three benches.
# (Keep gpurun_out small: more than 64 MiB is not copied back.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
FULL="--set full --clock-control none"
LM="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"
B="--steps 1 --warmup 3 --no-inference --no-cpu-baseline --no-torch-gpu"
timeout 900 ncu $LM -s 1400 -c 1000 --csv --log-file $O/r02_launches_vitl224_v2.csv python bench.py $B > $O/r02_ncu_l224b.log 2>&1
timeout 900 ncu $LM -s 1400 -c 1000 --csv --log-file $O/r02_launches_vitl384_v2.csv python bench.py --workload vitl384 $B > $O/r02_ncu_l384b.log 2>&1
timeout 600 ncu $FULL --import-source on -k regex:attn_.*_long_kernel -s 4 -c 2 -o $O/r02_attn_long_v2 -f python scripts/gpu_attn_prof.py 128 577 16 > $O/r02_ncu_attn_long2.log 2>&1
timeout 600 ncu $FULL --import-source on -k regex:attn_.*_fused_kernel -s 6 -c 2 -o $O/r02_attn_fused_v6 -f python scripts/gpu_attn_prof.py 256 197 16 > $O/r02_ncu_attn_fused6.log 2>&1
timeout 600 ncu $FULL -k regex:gemm_bf16_tcgen05_kernel -s 1 -c 12 -o $O/r02_gemm_variants -f python scripts/gpu_gemm_variants_once.py > $O/r02_ncu_gemm_variants.log 2>&1
ncu -i $O/r02_gemm_variants.ncu-rep --page raw --csv > $O/r02_gemm_variants_raw.csv 2>/dev/null && rm -f $O/r02_gemm_variants.ncu-rep
timeout 300 python bench.py --workload vitb224 --no-cpu-baseline --no-torch-gpu > $O/r02_bench_vitb224.json 2> $O/r02_bench_vitb224.err
timeout 300 python bench.py --workload vitl384 --no-cpu-baseline --no-torch-gpu > $O/r02_bench_vitl384_c.json 2> $O/r02_bench_vitl384_c.err
timeout 600 python bench.py > $O/r02_bench_f.json 2> $O/r02_bench_f.err
du -sh $O; ls -la $O
