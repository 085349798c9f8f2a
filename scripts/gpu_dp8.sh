N=8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-inference --no-cpu-baseline "$@" 2>gpurun_out/err8_$1_$2.log | tail -1 > gpurun_out/bench8_$2.json; python -c "import sys,json; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2:], round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['e2e']['value'], d.get('e2e_augmented',{}).get('value'))" gpurun_out/bench8_$2.json "$@"; }
run --bucket-mb 96
run --bucket-mb 100000
