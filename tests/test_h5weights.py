"""CPU: Keras-HDF5 weight reader/writer (no h5py in this image, so this is a self-consistency
round trip plus structural checks against the HDF5 file-format spec constants)."""
import os
import struct

import numpy as np
import pytest

import synth


def test_keras_h5_round_trip(tmp_path):
    from mrcnn import h5weights
    w = synth.make_random_weights(3, 4)
    small = {k: w[k] for k in ("conv1", "bn_conv1", "res2a_branch2a", "bn2a_branch2a", "rpn_conv_shared", "rpn_class_raw",
                               "rpn_bbox_pred", "mrcnn_class_logits", "mrcnn_mask_deconv", "mrcnn_mask")}
    path = str(tmp_path / "mask_rcnn_test_0001.h5")
    h5weights.write_keras_weights(path, small)
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0            # superblock v0
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)              # end-of-file address
    back = h5weights.read_keras_weights(path)
    assert set(back) == set(small)                                       # rpn_model was un-nested by name
    for k in small:
        assert len(back[k]) == len(small[k])
        for a, b in zip(back[k], small[k]):
            assert a.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)
    f = h5weights.H5File(path)
    names = [h5weights._as_str(n) for n in f.attributes(f.root["ohdr"])["layer_names"]]
    assert "rpn_model" in names and "rpn_conv_shared" not in names       # Keras nests the RPN as one layer
    assert h5weights._as_str(f.attributes(f.root["ohdr"])["keras_version"]) == "2.2.4"
    g = f.group_links(f.root["ohdr"])["rpn_model"]
    wn = [h5weights._as_str(n) for n in f.attributes(g)["weight_names"]]
    assert wn == ["rpn_conv_shared/kernel:0", "rpn_conv_shared/bias:0", "rpn_class_raw/kernel:0", "rpn_class_raw/bias:0",
                  "rpn_bbox_pred/kernel:0", "rpn_bbox_pred/bias:0"]


def test_full_model_file_layout_and_many_layers(tmp_path):
    from mrcnn import h5weights
    w = synth.make_random_weights(1, 4)
    # all 235 layers, but tiny arrays so the file stays small: exercises multi-SNOD groups
    tiny = {k: [np.asarray(a).ravel()[:7].astype(np.float32) for a in v] for k, v in w.items()}
    path = str(tmp_path / "all_layers.h5")
    h5weights.write_keras_weights(path, tiny)
    back = h5weights.read_keras_weights(path)
    assert set(back) == set(tiny) and len(back) == 235
    for k in tiny:
        for a, b in zip(back[k], tiny[k]):
            assert np.array_equal(a, b)


def test_not_hdf5(tmp_path):
    from mrcnn import h5weights
    p = tmp_path / "x.h5"
    p.write_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:2e22\nsize 255901152\n")
    with pytest.raises(IOError, match="not an HDF5 file"):
        h5weights.read_keras_weights(str(p))
