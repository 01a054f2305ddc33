#!/bin/bash
# round-2 call C: proposal popper tests, full GPU tests, bench with proposal clocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph_layers.py -q -m gpu -x -k proposal > gpurun_out/prop_tests.log 2>&1; echo "proposal tests exit $?"; tail -5 gpurun_out/prop_tests.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; grep -v "Invalid det bbox" gpurun_out/tests.log | tail -4
MRCNN_B200_PROPOSAL_CLOCKS=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "proposal phases" | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print({k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
print({k: round(v["ms_per_step"],4) for k,v in d["kernel_families"].items()})
PY
