#!/bin/bash
# One node, 2 GPUs: the tiled CLI run under torchrun (one rank per GPU, gloo for the catalogue gather) must produce
# the same set of sources as the single-process run.  Usage (on a box with >= 2 GPUs): tools/tile_multi_check.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/tile_multi
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import synth
from test_gpu_sfinder import write_fits
base = synth.radio_maps(16, 256)
rows = [np.concatenate([np.roll(base[(r * 4 + c) % 16], (r * 13, c * 29), axis=(0, 1)) for c in range(4)], axis=1) for r in range(4)]
write_fits("/tmp/tile_multi/mosaic.fits", np.concatenate(rows, axis=0)[:1000, :900])
PY
ARGS="detect --image /tmp/tile_multi/mosaic.fits --random_weights 0 --scoreThr 0.45 --nimg_per_gpu 4 --split_img_in_tiles --tile_xsize 256 --tile_ysize 256 --tile_xstep 0.75 --tile_ystep 1.0"
python caesar-mrcnn_b200/scripts/run.py $ARGS --detect_outfile_json /tmp/tile_multi/single.json > gpurun_out/tile_single.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    caesar-mrcnn_b200/scripts/run.py $ARGS --detect_outfile_json /tmp/tile_multi/multi.json > gpurun_out/tile_multi.log 2>&1
python - <<'PY'
import json
def key(s):
    # a merged source takes class / score from the LAST member of its group (reference quirk), which depends on the
    # order the tiles reach the master, i.e. on the number of ranks: class is compared for unmerged sources only
    cls = -1 if s["merged"] else int(s["class_id"])
    return (int(s["x1"]), int(s["x2"]), int(s["y1"]), int(s["y2"]), cls, len(s["pixels"]), bool(s["merged"]), bool(s["edge"]))
a = json.load(open("/tmp/tile_multi/single.json"))["sources"]
b = json.load(open("/tmp/tile_multi/multi.json"))["sources"]
ka, kb = sorted(map(key, a)), sorted(map(key, b))
if ka != kb:
    only_a = [k for k in ka if k not in kb][:5]
    only_b = [k for k in kb if k not in ka][:5]
    print("MISMATCH: %d vs %d sources; only single: %s; only multi: %s" % (len(a), len(b), only_a, only_b))
    raise SystemExit(1)
assert len(a) > 0
print("tile_multi_check ok: %d sources (%d merged across tiles, %d at edges) identical for 1 process and 2 ranks"
      % (len(a), sum(s["merged"] for s in a), sum(s["edge"] for s in a)))
PY
