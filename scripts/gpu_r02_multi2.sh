#!/bin/bash
# 8-GPU line of the 577-token workload (BASELINE config 5) + the two-device tests.
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload vitl384 > $O/r02_bench_${N}gpu_vitl384.json 2> $O/r02_bench_${N}gpu_vitl384.err
timeout 300 python -m pytest tests/test_gpu_dp.py -q -x > $O/r02_pytest_dp_${N}gpu.log 2>&1; tail -3 $O/r02_pytest_dp_${N}gpu.log
timeout 500 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_${N}gpu_vitl224_b.json 2> $O/r02_bench_${N}gpu_vitl224_b.err
python - <<PY
import json
for w in ("vitl384", "vitl224_b"):
    try:
        d = json.loads(open("$O/r02_bench_${N}gpu_%s.json" % w).read().strip().splitlines()[-1])
        print(w, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", d["clocks"], "identical", d.get("ranks_bit_identical_parameters"),
              "e2e", round(d["e2e"]["value"], 1), "infer", (d.get("inference") or {}).get("batches"), ((d.get("inference") or {}).get("e2e") or {}).get("batches"))
    except Exception as e:
        print(w, "FAILED", e)
PY
