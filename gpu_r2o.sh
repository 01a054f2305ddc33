#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "packed or async or end_to_end" 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value %.1f e2e %.1f packed %.1f ms %.3f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["e2e_packed_masks"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print({k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
PY
