#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -1 gpurun_out/bench_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --image-size 1024 --batch 8 > gpurun_out/bench_1024.log 2> gpurun_out/bench_1024.err; echo "bench1024 exit $?"; tail -2 gpurun_out/bench_1024.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1024.log').read().strip().splitlines()[-1])
print("S=1024 B=8", {k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'])
print(d['stage_ms_per_step'])
PY
