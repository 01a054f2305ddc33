#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}, d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/bench_n$N.err
