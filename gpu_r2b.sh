#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; grep -v "Invalid det bbox" gpurun_out/tests.log | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms %.3f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print({k: round(v,3) for k,v in d["stage_ms_per_step"].items()})
print({k: round(v["ms_per_step"],4) for k,v in d["kernel_families"].items()})
PY
MRCNN_B200_ROIALIGN_ROWS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows=0:', round(d['ms_per_step'],3), {k: round(v['ms_per_step'],4) for k,v in d['kernel_families'].items() if k=='roialign'})"
for g in 0 1; do MRCNN_B200_NATIVE_GRAPH=$g timeout 600 python tools/catalog_bench.py > gpurun_out/catalog_g$g.log 2>&1; echo "catalog native_graph=$g:"; tail -3 gpurun_out/catalog_g$g.log | cut -c1-600; done
