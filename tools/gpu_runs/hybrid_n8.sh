#!/bin/bash
# usage: hybrid_n8.sh N  (gpurun --gpus N): e2e of the detect bench with the dense share off and adaptive
N=${1:-8}
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" | head -8
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 30 --warmup 6 > gpurun_out/hyb_n${N}_$1.log 2> gpurun_out/hyb_n${N}_$1.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/hyb_n${N}_$1.log').read().strip().splitlines()[-1])
    print('$1 N=$N value %.0f e2e %.0f ms %.2f e2e_ms %.2f share %s packed %.0f threads %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('dense_share_images'), d['e2e_packed_masks']['value'], d['config']['host_threads']))
except Exception as e:
    print('no json', e); print(open('gpurun_out/hyb_n${N}_$1.err').read()[-1500:])
PY
}
MRCNN_B200_DENSE_SHARE=0 run off 29521
MRCNN_B200_DENSE_SHARE=auto run adaptive 29522
