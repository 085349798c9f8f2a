#!/bin/bash
# Round-2 multi-GPU evidence (run with gpurun --gpus N): data-parallel parity, the N-GPU bench of the headline and of the
# 577-token workload (each with the per-rank-replica inference leg), and the two-device tests.
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 scripts/gpu_dp_parity.py > $O/r02_dp_parity_${N}gpu.log 2>&1; echo "parity rc $?" >> $O/r02_dp_parity_${N}gpu.log
tail -5 $O/r02_dp_parity_${N}gpu.log
timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_${N}gpu_vitl224.json 2> $O/r02_bench_${N}gpu_vitl224.err
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload vitl384 > $O/r02_bench_${N}gpu_vitl384.json 2> $O/r02_bench_${N}gpu_vitl384.err
timeout 300 python -m pytest tests/test_gpu_dp.py -q -x > $O/r02_pytest_dp_${N}gpu.log 2>&1; tail -3 $O/r02_pytest_dp_${N}gpu.log
python - <<PY
import json
for w in ("vitl224", "vitl384"):
    try:
        d = json.loads(open("$O/r02_bench_${N}gpu_%s.json" % w).read().strip().splitlines()[-1])
        print(w, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", d["clocks"], "identical", d.get("ranks_bit_identical_parameters"),
              "e2e", round(d["e2e"]["value"], 1), "infer", (d.get("inference") or {}).get("batches"), ((d.get("inference") or {}).get("e2e") or {}).get("batches"))
    except Exception as e:
        print(w, "FAILED", e)
PY
