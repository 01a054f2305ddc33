"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-CPU fp32 / one small C++ file) of the caesar-mrcnn Mask R-CNN
*detect* hot path.  It exists to CHECK the CUDA path; it is never the
thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under ``caesar-mrcnn_b200/``
imports it, and the product path fails loudly when its CUDA library is missing.

Parity status
-------------
* The arithmetic of this path lives in third-party packages that are NOT under /root/reference and
  cannot be installed in this image: tensorflow==1.13.2, keras==2.2.4 (requirements.txt:19-20),
  astropy (requirements.txt:4), scikit-image<=0.15 (setup.py:45).  Their published algorithms are
  restated here.  The reference ships no tests, golden vectors or KATs for
  this path, so for those pieces the oracle is **parity unpinned** (cross-checked only against
  independent implementations: torchvision.ops.nms, cv2.resize interior pixels, torch conv2d).
* The pure-numpy pieces of the reference (anchors, norm/denorm_boxes, compose_image_meta,
  mold_image, resize_image bookkeeping, unmold_detections box maths) ARE pinned: the real reference
  functions were executed in the build container with the missing third-party imports stubbed
  (tests/golden/make_golden_from_reference.py) and their outputs are committed under tests/golden/.

Conventions that make CPU and GPU agree bit-for-bit on the index-producing stages
---------------------------------------------------------------------------------
* float32 everywhere, one rounding per reference op, no FMA contraction;
* exp/log are "evaluate in float64, round once to float32" (TF's Eigen pexp/plog cannot be
  reproduced without TF);
* round = half-to-even; float->int32 casts truncate, -inf -> INT_MIN.
"""
