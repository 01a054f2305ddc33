#!/bin/bash
mkdir -p gpurun_out
MRCNN_B200_GRAPH=1 timeout 1200 python -m pytest tests/test_gpu_engine.py -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
cat > /tmp/graph_probe.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "caesar-mrcnn_b200"))
import numpy as np, torch, synth
from mrcnn import model as modellib
from mrcnn.config import Config
B, S = 64, 256
class C(Config):
    NAME = "probe"; GPU_COUNT = 1; IMAGES_PER_GPU = B; NUM_CLASSES = 4; IMAGE_MIN_DIM = S; IMAGE_MAX_DIM = S
    RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64); MEAN_PIXEL = np.array([0, 0, 0]); DETECTION_MIN_CONFIDENCE = 0
m = modellib.MaskRCNN("inference", C(), "/tmp/x"); m.set_weights(synth.make_random_weights(0, 4))
maps = torch.from_numpy(synth.radio_maps(B, S)).cuda()
def run(n):
    for i in range(n): m.detect_maps(maps, device_only=True, _async=True)
    m.wait()
best = 1e9
for rep in range(4):
    run(4); torch.cuda.synchronize(); t = time.perf_counter(); run(30); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t) / 30 * 1e3)
print("GRAPH=%s: %.3f ms/batch (best of 4 x 30)" % (os.environ.get("MRCNN_B200_GRAPH", "0"), best))
PY
for g in 0 1 0 1; do MRCNN_B200_GRAPH=$g python /tmp/graph_probe.py 2>&1 | tail -1; done
