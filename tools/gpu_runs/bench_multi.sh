#!/bin/bash
# usage: gpu_multi2.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
PORT=29511
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "detect bench N=$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
    print('detect N=$N value %.0f e2e %.0f ms %.2f e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_n$N.err').read()[-1500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --mode train --gpus $N --steps 20 --warmup 3 > gpurun_out/train_bench_n$N.log 2> gpurun_out/train_bench_n$N.err; echo "train bench N=$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_n$N.log').read().strip().splitlines()[-1])
    print('train N=$N value %.1f img/s ms %.2f e2e_ms %.2f graph %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['cuda_graph']), d['allreduce'], d['phase_ms_per_step_eager'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/train_bench_n$N.err').read()[-2500:])
PY
grep -i "warn\|error" gpurun_out/train_bench_n$N.err | head -5
