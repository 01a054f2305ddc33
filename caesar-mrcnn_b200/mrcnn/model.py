"""mrcnn.model — B200-native `MaskRCNN` with the reference's inference API surface
(mrcnn/model.py:1911-2884): MaskRCNN(mode, config, model_dir), load_weights(filepath, by_name,
exclude), detect(images, verbose), detect_molded, mold_inputs, unmold_detections, get_anchors,
print_model, plus the module functions compose_image_meta / parse_image_meta / mold_image /
unmold_image / compute_backbone_shapes.

`keras_model.predict` is replaced by the C++/CUDA engine of libmrcnn_b200.so (csrc/engine.cu): one
static plan of sm_100a kernels per (BATCH_SIZE, IMAGE_SHAPE).  mode='training' is not part of this
build and raises.  No TensorFlow, no Keras, no CPU fallback.
"""
import ctypes
import datetime
import errno
import logging
import os
import re
import time

import numpy as np

from . import _native, utils
from .utils import compute_backbone_shapes  # noqa: F401  (re-exported, as in the reference)

logger = logging.getLogger("mrcnn")

# names of the graph outputs in keras_model.predict order (mrcnn/model.py:2156-2158)
OUTPUT_NAMES = ["detections", "mrcnn_class", "mrcnn_bbox", "mrcnn_mask", "rpn_rois", "rpn_class", "rpn_bbox"]


def log(text, array=None):
    if array is not None:
        text = text.ljust(25) + "shape: {:20}  ".format(str(array.shape))
        if array.size:
            text += "min: {:10.5f}  max: {:10.5f}".format(array.min(), array.max())
        text += "  {}".format(array.dtype)
    print(text)


# --------------------------------------------------------------------------------------------
# data formatting (mrcnn/model.py:2891-2975)
# --------------------------------------------------------------------------------------------

def compose_image_meta(image_id, original_image_shape, image_shape, window, scale, active_class_ids):
    return np.array([image_id] + list(original_image_shape) + list(image_shape) + list(window) + [scale]
                    + list(active_class_ids))


def parse_image_meta(meta):
    return {
        "image_id": meta[:, 0].astype(np.int32),
        "original_image_shape": meta[:, 1:4].astype(np.int32),
        "image_shape": meta[:, 4:7].astype(np.int32),
        "window": meta[:, 7:11].astype(np.int32),
        "scale": meta[:, 11].astype(np.float32),
        "active_class_ids": meta[:, 12:].astype(np.int32),
    }


def mold_image(images, config):
    return images.astype(np.float32) - config.MEAN_PIXEL


def expand_mask_bits(mask_bits, mask_shape):
    """One image's packed detection masks (result(expand=False): uint32 [H*W, words]) -> the reference's [H, W, N] bool array
    (mrcnn/model.py:2613-2619); same host routine as the default path, one image at a time."""
    H, W, n = mask_shape
    if n == 0:
        return np.empty((H, W, 0))
    bits = np.ascontiguousarray(mask_bits).view(np.uint32)
    dense = np.empty((H * W * n,), dtype=np.uint8)
    cnt = np.array([n], dtype=np.int32)
    dst = (ctypes.c_void_p * 1)(dense.ctypes.data)
    _native.check(_native.lib().mrcnn_host_expand_mask_bits(bits.ctypes.data, 1, H * W, bits.shape[1], cnt.ctypes.data, dst, 0),
                  "host_expand_mask_bits")
    return dense.reshape(H, W, n).view(np.bool_)


def unmold_image(normalized_images, config):
    return (normalized_images + config.MEAN_PIXEL).astype(np.uint8)


def graph_limit_problems(config):
    """Hard limits of the fused sm_100a kernels, checked when the model is built (not at first predict):
    -> list of messages, empty when the configuration fits."""
    out = []
    if not 1 <= int(config.NUM_CLASSES) <= 6:
        out.append("NUM_CLASSES=%d: the fused class/bbox head packs 5*NUM_CLASSES outputs into one 32-wide tile (max 6 "
                   "classes including background)" % config.NUM_CLASSES)
    if not 1 <= len(config.RPN_ANCHOR_RATIOS) <= 5:
        out.append("len(RPN_ANCHOR_RATIOS)=%d: the fused RPN head packs 6*anchors_per_location outputs into one 32-wide "
                   "tile (max 5)" % len(config.RPN_ANCHOR_RATIOS))
    if int(config.POST_NMS_ROIS_INFERENCE) > 1024:
        out.append("POST_NMS_ROIS_INFERENCE=%d: DetectionLayer kernel handles at most 1024 ROIs per image" % config.POST_NMS_ROIS_INFERENCE)
    if int(config.DETECTION_MAX_INSTANCES) > 256:
        out.append("DETECTION_MAX_INSTANCES=%d: at most 256 detections per image" % config.DETECTION_MAX_INSTANCES)
    if int(config.TOP_DOWN_PYRAMID_SIZE) % 64 or int(config.FPN_CLASSIF_FC_LAYERS_SIZE) % 64:
        out.append("TOP_DOWN_PYRAMID_SIZE and FPN_CLASSIF_FC_LAYERS_SIZE must be multiples of 64")
    if list(config.BACKBONE_STRIDES) != [4, 8, 16, 32, 64]:
        out.append("BACKBONE_STRIDES must be [4, 8, 16, 32, 64]")
    return out


# training-path functions of the reference's mrcnn.model namespace (mrcnn/model.py:1277-1381, 1536-1644, 1721-1904)
from .training import build_rpn_targets, data_generator, load_image_gt  # noqa: E402,F401


# --------------------------------------------------------------------------------------------
# MaskRCNN
# --------------------------------------------------------------------------------------------

def _dense_share_requested():
    v = os.environ.get("MRCNN_B200_DENSE_SHARE", "").strip()
    return v != "" and v != "0"


class _PinnedSet(object):
    """One batch worth of host result buffers: rois, class ids, scores, counts and the pixel-major mask bits
    [B, H0*W0, DW] uint32 land in PINNED memory (device->host copies); `dense` is the plain host buffer the
    reference-contract [H0,W0,N] bool masks are expanded into (mrcnn_host_expand_mask_bits)."""

    def __init__(self, torch, B, D, H0, W0, dw, pin=True):
        # allocated pinned directly (tensor.pin_memory() would allocate pageable memory first and copy it)
        self.key = (B, D, H0, W0)
        self.dw = dw
        self.tensors = (torch.empty((B, D, 4), dtype=torch.int32, pin_memory=pin), torch.empty((B, D), dtype=torch.int32, pin_memory=pin),
                        torch.empty((B, D), dtype=torch.float32, pin_memory=pin), torch.empty((B,), dtype=torch.int32, pin_memory=pin),
                        torch.empty((B, H0 * W0, dw), dtype=torch.int32, pin_memory=pin))
        # pinned as well when the hybrid delivery is on (MRCNN_B200_DENSE_SHARE): the DMA engine then writes the dense
        # masks of some images straight into it
        self.dense = torch.empty((B, H0 * W0 * D), dtype=torch.uint8, pin_memory=pin and _dense_share_requested())


class _Lease(object):
    """Alive while any numpy array handed to the caller still views the set; returns it to the pool afterwards."""

    def __init__(self, pool, pset):
        self._pool, self._pset = pool, pset

    def __del__(self):
        try:
            self._pool.setdefault(self._pset.key, []).append(self._pset)
        except Exception:
            pass


class _LeasedArray(object):
    """numpy-convertible window on one host tensor; np.asarray(...) keeps this object (and so the lease) as .base"""

    def __init__(self, lease, tensor, typestr):
        self._lease, self._tensor = lease, tensor
        self.__array_interface__ = {"shape": tuple(tensor.shape), "typestr": typestr, "data": (tensor.data_ptr(), False), "version": 3}


def _lease_arrays(pool, pset):
    """-> (rois, class_ids, scores, counts, mask_bits, dense, (H0, W0)): numpy views sharing one lease"""
    lease = _Lease(pool, pset)
    types = ("<i4", "<i4", "<f4", "<i4", "<u4", "|u1")
    return tuple(np.asarray(_LeasedArray(lease, t, ts)) for t, ts in zip(pset.tensors + (pset.dense,), types)) + (pset.key[2:],)


class _PendingDetection(object):
    """Handle of an asynchronous detect_maps call (keeps the pinned buffers and inputs alive)."""

    def __init__(self, model, bufs, maps, slot, n_dense=0):
        self._model, self._bufs, self._maps, self._done, self._slot, self._n_dense = model, bufs, maps, None, slot, n_dense

    def result(self, expand=True):
        """expand=False (extension): leave each image's masks as the packed bits that crossed PCIe — dict keys rois,
        class_ids, scores, mask_bits (uint32 [H*W, words], bit k of word j = detection 32*j + k), mask_shape (H, W, N) —
        for consumers that do not need the dense [H,W,N] bool arrays at once (expand_mask_bits makes them on demand)."""
        if self._done is None:
            # first everything but the dense share; the host expands its images while the DMA copy of the share drains
            t0 = time.perf_counter()
            _native.check(self._model._lib.mrcnn_engine_wait_slot_packed(self._model._engine, self._slot), "engine_wait_slot_packed")
            wait_ms = (time.perf_counter() - t0) * 1e3
            self._done = self._model._results_from_buffers(self._bufs, expand, self._n_dense, self._model, self._slot, wait_ms)
            self._maps = None
        return self._done


class _PendingDeviceDetection(object):
    """Handle of detect_maps(..., masks_on_device=True): boxes / class ids / scores / counts arrive in small pinned
    buffers through the copy stream, the [B,H,W,D] uint8 masks stay in the engine's result slot on the device
    (valid until the second next detect call reuses the slot)."""

    def __init__(self, model, bufs, maps, slot, H0, W0):
        self._model, self._bufs, self._maps, self._slot, self._done = model, bufs, maps, slot, None
        self.frame_hw = (H0, W0)

    def result(self):
        """-> dict(rois [B,D,4], class_ids [B,D], scores [B,D], counts [B] (host), masks_ptr (device address), depth)"""
        if self._done is None:
            m = self._model
            _native.check(m._lib.mrcnn_engine_wait_slot(m._engine, self._slot), "engine_wait_slot")
            ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
            name = b"unmold_masks" if self._slot == 0 else b"unmold_masks#1"
            _native.check(m._lib.mrcnn_engine_tensor(m._engine, name, ctypes.byref(ptr), ctypes.byref(nbytes)), "unmold_masks")
            rois, cls, scores, counts = self._bufs[:4]
            self._done = {"rois": rois, "class_ids": cls, "scores": scores, "counts": counts, "masks_ptr": ptr.value,
                          "depth": m.config.DETECTION_MAX_INSTANCES}
            self._maps = None
        return self._done


class MaskRCNN(object):
    """Mask R-CNN inference model; drop-in for mrcnn.model.MaskRCNN(mode='inference')."""

    def __init__(self, mode, config, model_dir, device=None):
        assert mode in ['training', 'inference']
        self.mode = mode
        self.config = config
        self.model_dir = model_dir
        self.set_log_dir()
        self._engine = None
        self._weights_loaded = False
        self._device = device
        self._graph = None
        self.keras_model = self.build(mode=mode, config=config)

    # -- construction ---------------------------------------------------------------------------
    def build(self, mode, config):
        if mode == "training":
            # the training graph (mrcnn/model.py:2068-2132) lives in mrcnn/training.py; it owns its own parameters
            from . import training
            self._graph = training.TrainGraph(config, device=self._device)
            self._device = self._graph.device.index
            return self
        h, w = config.IMAGE_SHAPE[:2]
        if h / 2 ** 6 != int(h / 2 ** 6) or w / 2 ** 6 != int(w / 2 ** 6):
            raise Exception("Image size must be dividable by 2 at least 6 times "
                            "to avoid fractions when downscaling and upscaling."
                            "For example, use 256, 320, 384, 448, 512, ... etc. ")
        if h != w:
            raise NotImplementedError("mrcnn (B200 build): square IMAGE_SHAPE only")
        if config.BACKBONE != "resnet101":
            raise NotImplementedError("mrcnn (B200 build): BACKBONE must be 'resnet101' (SURVEY.md §8a row a17)")
        if config.IMAGE_CHANNEL_COUNT != 3:
            raise NotImplementedError("mrcnn (B200 build): IMAGE_CHANNEL_COUNT must be 3 (SURVEY.md §8a row a17)")
        problems = graph_limit_problems(config)
        if problems:
            raise NotImplementedError("mrcnn (B200 build): " + "; ".join(problems))
        torch = utils._torch()
        lib = _native.lib()
        if self._device is None:
            self._device = torch.cuda.current_device()
        cfg = _native.EngineConfig()
        cfg.batch_size = int(config.BATCH_SIZE)
        cfg.image_size = int(h)
        cfg.num_classes = int(config.NUM_CLASSES)
        cfg.pre_nms_limit = int(config.PRE_NMS_LIMIT)
        cfg.post_nms_rois = int(config.POST_NMS_ROIS_INFERENCE)
        cfg.detection_max_instances = int(config.DETECTION_MAX_INSTANCES)
        cfg.pool_size = int(config.POOL_SIZE)
        cfg.mask_pool_size = int(config.MASK_POOL_SIZE)
        cfg.fc_layers_size = int(config.FPN_CLASSIF_FC_LAYERS_SIZE)
        cfg.top_down_pyramid_size = int(config.TOP_DOWN_PYRAMID_SIZE)
        cfg.anchors_per_location = len(config.RPN_ANCHOR_RATIOS)
        cfg.rpn_nms_threshold = float(config.RPN_NMS_THRESHOLD)
        cfg.detection_min_confidence = float(config.DETECTION_MIN_CONFIDENCE or 0.0)
        cfg.detection_nms_threshold = float(config.DETECTION_NMS_THRESHOLD)
        for i in range(4):
            cfg.rpn_bbox_std_dev[i] = float(np.float32(config.RPN_BBOX_STD_DEV[i]))
            cfg.bbox_std_dev[i] = float(np.float32(config.BBOX_STD_DEV[i]))
        for i in range(5):
            cfg.backbone_strides[i] = int(config.BACKBONE_STRIDES[i])
        handle = ctypes.c_void_p()
        _native.check(lib.mrcnn_engine_create(ctypes.byref(cfg), int(self._device), ctypes.byref(handle)), "engine_create")
        self._engine = handle
        self._lib = lib
        # everything this model launches (torch plumbing included) is ordered on the engine's stream
        self._stream = torch.cuda.ExternalStream(lib.mrcnn_engine_stream(handle), device=int(self._device))
        self._pinned = {}
        self._result_pool = {}
        self._dense_share_state = None
        return self     # callers only use .predict() on it

    def __del__(self):
        try:
            if getattr(self, "_engine", None):
                self._lib.mrcnn_engine_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    def layer_table(self):
        """[(name, kind, [kernel shape])] of the weighted layers the inference graph expects."""
        lib, out = self._lib, []
        kinds = {0: "conv", 1: "bn", 2: "dense", 3: "deconv"}
        for i in range(lib.mrcnn_engine_num_layers(self._engine)):
            name = ctypes.c_char_p()
            kind, nw = ctypes.c_int(), ctypes.c_int()
            shape = (ctypes.c_int * 4)()
            _native.check(lib.mrcnn_engine_layer_info(self._engine, i, ctypes.byref(name), ctypes.byref(kind),
                                                      ctypes.byref(nw), shape))
            out.append((name.value.decode(), kinds[kind.value], [s for s in shape if s > 0], nw.value))
        return out

    def print_model(self):
        print("mask_rcnn (B200 engine): %d weighted layers" % self._lib.mrcnn_engine_num_layers(self._engine))
        for name, kind, shape, _ in self.layer_table():
            print("  {:28} {:7} {}".format(name, kind, shape))

    # -- weights --------------------------------------------------------------------------------
    def set_weights(self, weights, exclude=None, allow_missing=False):
        """weights: {layer_name: [arrays in Keras layer.weights order]}."""
        if self.mode == "training":
            self._graph.params.set_weights(weights, exclude=exclude)
            self._weights_loaded = True
            return
        if self._weights_loaded:
            raise RuntimeError("weights already loaded into this engine; create a new MaskRCNN")
        lib = self._lib
        known = {name for name, _, _, _ in self.layer_table()}
        for name, arrays in weights.items():
            if name not in known:
                continue                                    # by_name semantics: unknown layers are skipped
            if exclude and name in exclude:                 # NB: `in` on a str is a substring test,
                continue                                    # as in the reference (model.py:2229, run.py:1738)
            for wi, arr in enumerate(arrays):
                a = np.ascontiguousarray(arr, dtype=np.float32)
                _native.check(lib.mrcnn_engine_set_weight(self._engine, name.encode(), wi, a.ctypes.data, a.size),
                              "load_weights(%s)" % name)
        _native.check(lib.mrcnn_engine_finalize(self._engine, 1 if (allow_missing or exclude) else 0), "load_weights")
        self._weights_loaded = True
        self._set_anchors()

    def load_weights(self, filepath, by_name=False, exclude=None):
        """Keras-HDF5 weight file -> engine (reference: mrcnn/model.py:2197-2239). Loading is always
        by layer name (the inference graph has no positional layer list)."""
        from . import h5weights
        weights = h5weights.read_keras_weights(filepath)
        # make a mis-parsed or foreign file visible: how many of the graph's weighted layers the file provides
        known = [n for n, _, _ in self._graph.params.specs] if self.mode == "training" else [n for n, _, _, _ in self.layer_table()]
        found = [n for n in known if n in weights and not (exclude and n in exclude)]
        missing = [n for n in known if n not in weights]
        log("load_weights(%s): %d of %d weighted layers loaded, %d missing%s, %d entries of the file unused" % (
            os.path.basename(str(filepath)), len(found), len(known), len(missing),
            (" (" + ", ".join(missing[:6]) + (" ..." if len(missing) > 6 else "") + ")") if missing else "",
            len([n for n in weights if n not in known])))
        self.set_weights(weights, exclude=exclude)
        self.set_log_dir(filepath)

    def get_weights(self):
        """{layer_name: [arrays in Keras layer.weights order]} of the training graph."""
        assert self.mode == "training", "Create model in training mode."
        return self._graph.params.get_weights()

    def save_weights(self, filepath):
        """Keras `save_weights` layout (what ModelCheckpoint(save_weights_only=True) writes, mrcnn/model.py:2452-2453)."""
        from . import h5weights
        w = self.get_weights()
        h5weights.write_keras_weights(filepath, w, layer_order=[n for n, _, _ in self._graph.params.specs])

    def compile(self, learning_rate, momentum):
        """reference: mrcnn/model.py:2259-2330 — SGD(lr, momentum, clipnorm=GRADIENT_CLIP_NORM), the five weighted losses
        and the L2 regulariser; here it creates the Trainer (bucketed all-reduce + fused optimiser kernels)."""
        from . import training
        assert self.mode == "training", "Create model in training mode."
        self._trainer = training.Trainer(self._graph, learning_rate=learning_rate, momentum=momentum)
        return self._trainer

    def set_trainable(self, layer_regex, keras_model=None, indent=0, verbose=1):
        assert self.mode == "training", "Create model in training mode."
        self._graph.set_trainable(layer_regex)
        if verbose > 0:
            log("Selecting layers to train")
            for (name, role), t in self._graph.masters.items():
                if role == "kernel" and t.requires_grad:
                    log("{}{:20}".format(" " * indent, name))

    def train(self, train_dataset, val_dataset, learning_rate, epochs, layers, augmentation=None, custom_callbacks=None,
              no_augmentation_sources=None, n_worker_threads=-1, class_weights=None, draw_loss=False):
        """reference: mrcnn/model.py:2395-2517.  Same arguments; `epochs` is the total epoch count (earlier epochs of a
        resumed run count), STEPS_PER_EPOCH batches per epoch, VALIDATION_STEPS validation batches (losses only) and one
        weights file per epoch at `checkpoint_path` (rank 0).  Under torchrun every rank trains on its own generator and the
        gradients are averaged over NVLink (training.GradReducer) — the replacement of ParallelModel.  TensorBoard,
        Keras callbacks, multiprocessing loaders and the loss plot are not part of this build; returns the history dict."""
        from . import training
        assert self.mode == "training", "Create model in training mode."
        if augmentation is not None or custom_callbacks or class_weights is not None:
            raise NotImplementedError("mrcnn (B200 build): augmentation / custom_callbacks / class_weights")
        cfg, g = self.config, self._graph
        torch = g.torch
        layers = training.LAYER_REGEX.get(layers, layers)
        train_gen = training.data_generator(train_dataset, cfg, shuffle=True, batch_size=cfg.BATCH_SIZE,
                                            no_augmentation_sources=no_augmentation_sources)
        val_gen = training.data_generator(val_dataset, cfg, shuffle=True, batch_size=cfg.BATCH_SIZE) if val_dataset is not None else None
        log("\nStarting at epoch {}. LR={}\n".format(self.epoch, learning_rate))
        log("Checkpoint Path: {}".format(self.checkpoint_path))
        self.set_trainable(layers, verbose=0)
        trainer = self.compile(learning_rate, cfg.LEARNING_MOMENTUM)
        trainer.broadcast_parameters()
        rank = trainer.reducer.dist.get_rank() if trainer.reducer.world > 1 else 0
        history = {"loss": [], "val_loss": []}
        for n in training.LOSS_NAMES:
            history[n] = []
            history["val_" + n] = []
        captured = False
        for epoch in range(self.epoch, epochs):
            sums = {}
            for _ in range(int(cfg.STEPS_PER_EPOCH)):
                inputs, _unused = next(train_gen)
                dev = g.to_device(inputs)
                if not captured and os.environ.get("MRCNN_B200_TRAIN_GRAPH", "1") != "0":
                    trainer.capture(dev)           # the whole step as one CUDA graph from here on (eager if that fails)
                captured = True
                ls = trainer.train_step(dev)
                for k, v in ls.items():
                    sums[k] = sums.get(k, 0.0) + float(v)
            for k, v in sums.items():
                history[k].append(v / max(1, int(cfg.STEPS_PER_EPOCH)))
            if val_gen is not None and int(cfg.VALIDATION_STEPS) > 0:
                vs = {}
                with torch.no_grad():
                    for _ in range(int(cfg.VALIDATION_STEPS)):
                        inputs, _unused = next(val_gen)
                        total, ls = g.forward(g.to_device(inputs))
                        ls = dict(ls)
                        ls["loss"] = total
                        for k, v in ls.items():
                            vs[k] = vs.get(k, 0.0) + float(v)
                for k, v in vs.items():
                    history["val_" + k].append(v / int(cfg.VALIDATION_STEPS))
            log("Epoch {}/{}: ".format(epoch + 1, epochs) + " ".join("{}={:.4f}".format(k, history[k][-1]) for k in history if history[k]))
            if rank == 0:
                os.makedirs(self.log_dir, exist_ok=True)
                self.save_weights(self.checkpoint_path.format(epoch=epoch + 1))
        self.epoch = max(self.epoch, epochs)
        return history

    def find_last(self):
        dir_names = next(os.walk(self.model_dir))[1]
        key = self.config.NAME.lower()
        dir_names = sorted(filter(lambda f: f.startswith(key), dir_names))
        if not dir_names:
            raise FileNotFoundError(errno.ENOENT, "Could not find model directory under {}".format(self.model_dir))
        dir_name = os.path.join(self.model_dir, dir_names[-1])
        checkpoints = sorted(filter(lambda f: f.startswith("mask_rcnn"), next(os.walk(dir_name))[2]))
        if not checkpoints:
            raise FileNotFoundError(errno.ENOENT, "Could not find weight files in {}".format(dir_name))
        return os.path.join(dir_name, checkpoints[-1])

    def set_log_dir(self, model_path=None):
        """log_dir / checkpoint_path / epoch bookkeeping (mrcnn/model.py:2357-2393); strings only."""
        self.epoch = 0
        now = datetime.datetime.now()
        if model_path:
            regex = r".*[/\\][\w-]+(\d{4})(\d{2})(\d{2})T(\d{2})(\d{2})[/\\]mask\_rcnn\_[\w-]+(\d{4})\.h5"
            m = re.match(regex, str(model_path))
            if m:
                now = datetime.datetime(int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)))
                self.epoch = int(m.group(6)) - 1 + 1
        name = (self.config.NAME or "mrcnn").lower()
        self.log_dir = os.path.join(self.model_dir or ".", "{}{:%Y%m%dT%H%M}".format(name, now))
        self.checkpoint_path = os.path.join(self.log_dir, "mask_rcnn_{}_*epoch*.h5".format(name)).replace("*epoch*", "{epoch:04d}")

    # -- anchors --------------------------------------------------------------------------------
    def get_anchors(self, image_shape):
        backbone_shapes = compute_backbone_shapes(self.config, image_shape)
        if not hasattr(self, "_anchor_cache"):
            self._anchor_cache = {}
        key = tuple(image_shape)
        if key not in self._anchor_cache:
            a = utils.generate_pyramid_anchors(self.config.RPN_ANCHOR_SCALES, self.config.RPN_ANCHOR_RATIOS, backbone_shapes,
                                               self.config.BACKBONE_STRIDES, self.config.RPN_ANCHOR_STRIDE)
            self.anchors = a
            self._anchor_cache[key] = utils.norm_boxes(a, image_shape[:2])
        return self._anchor_cache[key]

    def _set_anchors(self):
        a = np.ascontiguousarray(self.get_anchors(tuple(int(v) for v in self.config.IMAGE_SHAPE)), dtype=np.float32)
        _native.check(self._lib.mrcnn_engine_set_anchors(self._engine, a.ctypes.data, a.shape[0]), "set_anchors")

    # -- molding --------------------------------------------------------------------------------
    def _mold_inputs_device(self, images):
        """-> (molded CUDA float32 [B,S,S,3], image_metas float64 [B,12+NC], windows int [B,4])."""
        torch = utils._torch()
        cfg = self.config
        S = int(cfg.IMAGE_SHAPE[0])
        molded = torch.empty((len(images), S, S, 3), dtype=torch.float32, device="cuda:%d" % self._device)
        metas, windows = [], []
        groups = {}
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise NotImplementedError("mrcnn (B200 build): detect() takes uint8 [H,W,3] images (SURVEY.md §8a row a17)")
            groups.setdefault(im.shape, []).append(i)
        geo = {}
        for shape, idxs in groups.items():
            h, w = shape[:2]
            scale, out_hw, top_left, window, _pad = utils.square_geometry(h, w, cfg.IMAGE_MIN_DIM, cfg.IMAGE_MAX_DIM,
                                                                          cfg.IMAGE_MIN_SCALE, cfg.IMAGE_RESIZE_MODE)
            if cfg.IMAGE_RESIZE_MODE == "none":
                assert (h, w) == (S, S), "After resizing, all images must have the same size. Check IMAGE_RESIZE_MODE and image sizes."
            geo[shape] = (scale, window)
            host = torch.from_numpy(np.stack([images[i] for i in idxs]))
            rgb = host.to(molded.device, non_blocking=False)
            sub = utils.mold_rgb8_device(rgb, None, out_hw, S, top_left, cfg.MEAN_PIXEL)
            molded[torch.as_tensor(idxs, device=molded.device)] = sub
        for im in images:
            scale, window = geo[im.shape]
            metas.append(compose_image_meta(0, im.shape, (S, S, 3), window, scale, np.zeros([cfg.NUM_CLASSES], dtype=np.int32)))
            windows.append(window)
        return molded, np.stack(metas), np.stack(windows)

    def mold_inputs(self, images):
        """reference signature (mrcnn/model.py:2519-2556): numpy (molded_images, image_metas, windows)."""
        torch = utils._torch()
        with torch.cuda.stream(self._stream):
            molded, metas, windows = self._mold_inputs_device(images)
            out = molded.cpu().numpy()
        return out, metas, windows

    # -- the graph ------------------------------------------------------------------------------
    def predict(self, inputs, verbose=0):
        """keras_model.predict([molded_images, image_metas, anchors]) -> the 7 graph outputs as numpy."""
        molded, metas = inputs[0], inputs[1]
        self._predict_device(molded, metas)
        return [self.read_tensor(n) for n in OUTPUT_NAMES]

    def _predict_device(self, molded, metas):
        if not self._weights_loaded:
            raise RuntimeError("load_weights() / set_weights() must be called before predict/detect")
        torch = utils._torch()
        B = self.config.BATCH_SIZE
        metas32 = np.ascontiguousarray(metas, dtype=np.float32)
        assert metas32.shape == (B, self.config.IMAGE_META_SIZE)
        with torch.cuda.stream(self._stream):
            d_meta = torch.from_numpy(metas32).to("cuda:%d" % self._device)
            if isinstance(molded, np.ndarray):
                molded = torch.from_numpy(np.ascontiguousarray(molded, dtype=np.float32)).to("cuda:%d" % self._device)
            assert tuple(molded.shape) == (B,) + tuple(int(v) for v in self.config.IMAGE_SHAPE), "molded images do not match IMAGE_SHAPE"
            _native.check(self._lib.mrcnn_engine_predict(self._engine, _native.ptr(molded.contiguous()), _native.ptr(d_meta), 0, 0),
                          "predict")

    _TENSOR_SHAPES = None

    def read_tensor(self, name):
        """Host copy (float32 / int32) of a named engine tensor (graph outputs and taps)."""
        c = self.config
        B, R, D, NC = c.BATCH_SIZE, c.POST_NMS_ROIS_INFERENCE, c.DETECTION_MAX_INSTANCES, c.NUM_CLASSES
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        _native.check(self._lib.mrcnn_engine_tensor(self._engine, name.encode(), ctypes.byref(ptr), ctypes.byref(nbytes)), name)
        int_names = ("topk_idx", "keep_idx", "keep_count", "roi_levels")
        dt = np.int32 if name in int_names or name.startswith("unmold_rois") or name in ("unmold_class_ids", "unmold_counts") else np.float32
        f32_names = set(OUTPUT_NAMES) | {"anchors", "input_image", "input_image_meta", "unmold_scores", "mrcnn_class_head_raw",
                                          "mrcnn_mask_logits"}
        is_wide = name in f32_names or dt == np.int32 or name.startswith("rpn_head_p")
        n = nbytes.value // 4 if is_wide else nbytes.value // 2
        out = np.empty((n,), dtype=dt)
        _native.check(self._lib.mrcnn_engine_read(self._engine, name.encode(), out.ctypes.data, out.nbytes), "read " + name)
        shapes = {"detections": (B, D, 6), "mrcnn_class": (B, R, NC), "mrcnn_bbox": (B, R, NC, 4),
                  "mrcnn_mask": (B, D, c.MASK_SHAPE[0], c.MASK_SHAPE[1], NC), "rpn_rois": (B, R, 4),
                  "rpn_class": (B, -1, 2), "rpn_bbox": (B, -1, 4), "keep_idx": (B, R), "keep_count": (B,),
                  "topk_idx": (B, -1), "roi_levels": (B, R),
                  "pooled": (B, R, c.POOL_SIZE, c.POOL_SIZE, c.TOP_DOWN_PYRAMID_SIZE),
                  "pooled_mask": (B, D, c.MASK_POOL_SIZE, c.MASK_POOL_SIZE, c.TOP_DOWN_PYRAMID_SIZE)}
        if name in shapes:
            return out.reshape(shapes[name])
        if re.match(r"^[PC][2-6]$", name):
            ch = c.TOP_DOWN_PYRAMID_SIZE if name[0] == "P" else {"C2": 256, "C3": 512, "C4": 1024, "C5": 2048}[name]
            side = int(round((n // (B * ch)) ** 0.5))
            return out.reshape(B, side, side, ch)
        return out

    def write_tensor(self, name, array):
        a = np.ascontiguousarray(array, dtype=np.int32 if array.dtype.kind in "iu" else np.float32)
        _native.check(self._lib.mrcnn_engine_write(self._engine, name.encode(), a.ctypes.data, a.nbytes), "write " + name)

    def run_stage(self, stage):
        _native.check(self._lib.mrcnn_engine_run_stage(self._engine, stage.encode()), "run_stage " + stage)

    def stage_times(self):
        names = (ctypes.c_char_p * 32)()
        ms = (ctypes.c_float * 32)()
        n = self._lib.mrcnn_engine_stage_times(self._engine, 32, names, ms)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    # -- detection ------------------------------------------------------------------------------
    def _result_buffers(self, H0, W0):
        """Host buffers for one batch of results, as numpy arrays (pinned: boxes / ids / scores / counts / mask bits;
        plain: the dense mask buffer). The arrays handed back to the caller are views of them; a set returns to this
        model's pool when the last such view is garbage-collected (page-faulting in a fresh 420 MB buffer costs more
        than a whole detect step, so sets are recycled, never freed)."""
        torch = utils._torch()
        c = self.config
        key = (c.BATCH_SIZE, c.DETECTION_MAX_INSTANCES, int(H0), int(W0))
        free = self._result_pool.setdefault(key, [])
        pset = free.pop() if free else _PinnedSet(torch, *key, dw=self._lib.mrcnn_mask_bits_words(c.DETECTION_MAX_INSTANCES))
        return _lease_arrays(self._result_pool, pset)

    # -- hybrid delivery of the dense masks (opt-in: MRCNN_B200_DENSE_SHARE=<k> | auto) -----------------------------
    # The reference's result contract is one dense [H, W, N] bool array per image: 6.5 MB per 256x256 image with 100
    # detections, 420 MB per 64-image batch, expanded from the packed bits by this rank's host cores.  On a host with few
    # cores per GPU the engine can instead expand the masks of the first k images on the device and let the DMA engine
    # write them into the (then pinned) dense buffer while the cores expand the rest.  `auto` lets k follow where a
    # pipelined caller spends its time (batch already on the host when result() asks: grow; waiting for the device or
    # the DMA copy: shrink).  Measured: with ONE expansion thread on a 1-GPU box 25.3 -> 13.4 ms per step;
    # on the 8-GPU box of this pool nothing is gained — its 133 GB/s of host-memory write bandwidth is the limit
    # whichever engine writes the 8 x 420 MB — so the default is off.
    def _dense_share(self, B, D):
        if self._dense_share_state is None:
            env = os.environ.get("MRCNN_B200_DENSE_SHARE", "").strip().lower()
            if env == "auto" and D % 4 == 0:
                self._dense_share_state = {"fixed": False, "k": float(B // 4)}
            elif env.isdigit() and D % 4 == 0:
                self._dense_share_state = {"fixed": True, "k": float(max(0, min(B, int(env))))}
            else:
                self._dense_share_state = {"fixed": True, "k": 0.0}
        return int(round(self._dense_share_state["k"]))

    def _balance_dense_share(self, B, n_dense, cpu_ms, slot, wait_ms=None, wait_dense_ms=0.0):
        st = self._dense_share_state
        if st is None or st["fixed"] or wait_ms is None:
            return
        if wait_ms > 0.5 or wait_dense_ms > 0.3:      # the device / DMA side is what this rank waits for
            k = st["k"] - 1.0
        else:                                         # everything was there already: the host's cores are the bottleneck
            k = st["k"] + 2.0
        st["k"] = float(min(max(0, B - 4), max(0.0, k)))
        st["last"] = {"wait_ms": float(wait_ms), "wait_dense_ms": float(wait_dense_ms), "cpu_ms": float(cpu_ms), "n_dense": n_dense}

    def reserve_result_buffers(self, count, H0, W0):
        """Pre-allocates `count` result sets for [H0, W0] frames (keeps the allocation and the first-touch page faults
        out of the first calls)."""
        sets = [self._result_buffers(H0, W0) for _ in range(count)]
        for bufs in sets:
            bufs[5].fill(0)
        del sets

    @staticmethod
    def _results_from_buffers(bufs, expand=True, n_dense=0, model=None, slot=0, wait_ms=None):
        """Expands the mask bits of one batch into the reference's [H0,W0,N] bool arrays (multi-threaded C++ behind the
        C ABI) and builds the detect()-style dicts (mrcnn/model.py:2697-2703). The first n_dense images arrived dense
        already (hybrid delivery, _dense_share): only the others are expanded here."""
        rois_n, cls_n, sc_n, cnt_n, bits_n, dense_n, (H0, W0) = bufs
        B, D = cls_n.shape
        npx, dw = bits_n.shape[1], bits_n.shape[2]
        if not expand:
            if model is not None and n_dense > 0:        # the dense share is still being written into this buffer set
                _native.check(model._lib.mrcnn_engine_wait_slot(model._engine, slot), "engine_wait_slot")
            return [{"rois": rois_n[i, :int(cnt_n[i])], "class_ids": cls_n[i, :int(cnt_n[i])], "scores": sc_n[i, :int(cnt_n[i])],
                     "mask_bits": bits_n[i], "mask_shape": (H0, W0, int(cnt_n[i]))} for i in range(B)]
        base = dense_n.ctypes.data
        rest = B - n_dense
        t0 = time.perf_counter()
        if rest > 0:
            dst = (ctypes.c_void_p * rest)(*[base + i * npx * D for i in range(n_dense, B)])
            _native.check(_native.lib().mrcnn_host_expand_mask_bits(bits_n[n_dense:].ctypes.data, rest, npx, dw,
                                                                    cnt_n[n_dense:].ctypes.data, dst, 0), "host_expand_mask_bits")
        if model is not None:
            cpu_ms = (time.perf_counter() - t0) * 1e3
            t1 = time.perf_counter()
            _native.check(model._lib.mrcnn_engine_wait_slot(model._engine, slot), "engine_wait_slot")     # the dense share
            model._balance_dense_share(B, n_dense, cpu_ms, slot, wait_ms, (time.perf_counter() - t1) * 1e3)
        out = []
        for i in range(B):
            n = int(cnt_n[i])
            # no detections: the reference returns np.empty(original_image_shape[:2] + (0,)) (float64), model.py:2618-2619
            masks = dense_n[i, :npx * n].reshape(H0, W0, n).view(np.bool_) if n > 0 else np.empty((H0, W0, 0))
            out.append({"rois": rois_n[i, :n], "class_ids": cls_n[i, :n], "scores": sc_n[i, :n], "masks": masks})
        return out

    def _detect_device(self, molded, metas, windows, orig_shapes):
        """engine predict + device unmold -> list of result dicts (numpy views of pinned buffers)."""
        torch = utils._torch()
        shapes = {tuple(s[:2]) for s in orig_shapes}
        if not self._weights_loaded:
            raise RuntimeError("load_weights() / set_weights() must be called before predict/detect")
        if len(shapes) != 1:
            # images of different original sizes in one batch (allowed by the reference as long as the molded shapes
            # agree, mrcnn/model.py:2655-2658): one graph pass, then the per-image GPU unmold
            if not molded.is_cuda:
                molded = molded.to("cuda:%d" % self._device)
            self._predict_device(molded, metas)
            det, masks = self.read_tensor("detections"), self.read_tensor("mrcnn_mask")
            out = []
            for i, shp in enumerate(orig_shapes):
                rois, cls, scores, full = self.unmold_detections(det[i], masks[i], shp, tuple(int(v) for v in self.config.IMAGE_SHAPE),
                                                                 windows[i])
                out.append({"rois": rois, "class_ids": cls, "scores": scores, "masks": full})
            return out
        H0, W0 = next(iter(shapes))
        metas32 = np.ascontiguousarray(metas, dtype=np.float32)
        wins = np.ascontiguousarray(windows, dtype=np.int32)
        bufs = self._result_buffers(H0, W0)
        orig = (ctypes.c_int * 2)(int(H0), int(W0))
        on_host = 0 if molded.is_cuda else 1
        with torch.cuda.stream(self._stream):
            _native.check(self._lib.mrcnn_engine_detect_molded(self._engine, _native.ptr(molded), on_host, metas32.ctypes.data, orig,
                                                               wins.ctypes.data, *[b.ctypes.data for b in bufs[:5]]), "detect")
        return self._results_from_buffers(bufs)

    def detect_maps_async(self, maps, zscale_contrasts=(0.25, 0.25, 0.25)):
        """Queues detect_maps and returns a handle whose .result() gives the detect()-style dicts. The
        device->host copy of this batch overlaps the compute of the next queued batch (at most two
        batches in flight; call .result() on the older one before queueing a third)."""
        return self.detect_maps(maps, zscale_contrasts, _async=True)

    def wait(self):
        _native.check(self._lib.mrcnn_engine_wait(self._engine), "engine_wait")

    def detect_maps(self, maps, zscale_contrasts=(0.25, 0.25, 0.25), device_only=False, _async=False, masks_on_device=False,
                    mask_format=None):
        """Fast path from FITS-like maps (extension; the numpy contract of detect() is unchanged):
        maps [BATCH_SIZE,H,W] float32 — numpy / pinned torch tensor (copied H2D) or a CUDA tensor —
        -> read_fits stretch + mold + graph + unmold in one C-ABI call. Returns detect()-style dicts,
        or None with device_only=True (results stay in the engine's 'unmold_*' tensors). masks_on_device=True queues
        the call and returns a _PendingDeviceDetection: only boxes / class ids / scores / counts are copied to the host,
        the full-frame masks stay in HBM for mrcnn.analyze (no [B,H,W,100] transfer).
        Host results travel as pixel-major mask bits (16 bytes per pixel for 100 detections) and are expanded to the
        reference's [H,W,N] bool arrays on the host; mask_format=0 with device_only leaves [B,H,W,D] uint8 masks in
        the engine tensor "unmold_masks" instead of the bits in "unmold_mask_bits"."""
        torch = utils._torch()
        c = self.config
        if not self._weights_loaded:
            raise RuntimeError("load_weights() / set_weights() must be called before predict/detect")
        if isinstance(maps, np.ndarray):
            maps = torch.from_numpy(np.ascontiguousarray(maps, dtype=np.float32))
        assert maps.dim() == 3 and maps.shape[0] == c.BATCH_SIZE and maps.dtype == torch.float32 and maps.is_contiguous()
        H0, W0 = int(maps.shape[1]), int(maps.shape[2])
        key = (H0, W0)
        if key not in self._pinned:
            scale, out_hw, top_left, window, _ = utils.square_geometry(H0, W0, c.IMAGE_MIN_DIM, c.IMAGE_MAX_DIM, c.IMAGE_MIN_SCALE,
                                                                       c.IMAGE_RESIZE_MODE)
            S = int(c.IMAGE_SHAPE[0])
            meta = compose_image_meta(0, (H0, W0, 3), (S, S, 3), window, scale, np.zeros([c.NUM_CLASSES], dtype=np.int32))
            metas32 = np.ascontiguousarray(np.stack([meta] * c.BATCH_SIZE), dtype=np.float32)
            wins = np.ascontiguousarray(np.stack([window] * c.BATCH_SIZE), dtype=np.int32)
            self._pinned[key] = (out_hw, top_left, metas32, wins)
        out_hw, top_left, metas32, wins = self._pinned[key]
        con = _native.float_array(list(zscale_contrasts))
        mean = _native.float_array([float(v) for v in np.asarray(c.MEAN_PIXEL).reshape(-1)[:3]])
        if masks_on_device:
            bufs = self._result_buffers(1, 1)                  # small pinned set: its 1x1 mask buffers are not used
            outs = [b.ctypes.data for b in bufs[:4]] + [None]
            _async = True
        else:
            bufs = None if device_only else self._result_buffers(H0, W0)
            outs = [None] * 5 if device_only else [b.ctypes.data for b in bufs[:5]]
        if mask_format is None:
            mask_format = 0 if masks_on_device else 1          # device consumers read bytes, host results travel as bits
        assert mask_format == 1 or masks_on_device or device_only, "host results are shipped as mask bits (mask_format=1)"
        slot = self._lib.mrcnn_engine_next_slot(self._engine)
        n_dense = 0
        if bufs is not None and not masks_on_device:
            n_dense = self._dense_share(c.BATCH_SIZE, c.DETECTION_MAX_INSTANCES)
            if n_dense > 0:
                _native.check(self._lib.mrcnn_engine_set_dense_output(self._engine, bufs[5].ctypes.data, n_dense), "set_dense_output")
        with torch.cuda.stream(self._stream):
            _native.check(self._lib.mrcnn_engine_detect_maps(self._engine, _native.ptr(maps), 0 if maps.is_cuda else 1, H0, W0, con, mean,
                                                             int(out_hw[0]), int(out_hw[1]), int(top_left[0]), int(top_left[1]),
                                                             metas32.ctypes.data, wins.ctypes.data, *outs, mask_format, 1 if _async else 0),
                          "detect_maps")
        if masks_on_device:
            return _PendingDeviceDetection(self, bufs, maps, slot, H0, W0)
        if device_only:
            return None          # with _async=True nothing has been waited for: call wait() before reading tensors
        if _async:
            return _PendingDetection(self, bufs, maps, slot, n_dense)
        return self._results_from_buffers(bufs, True, n_dense, self, 0)

    def kernel_times(self):
        """{family: (ms, launches)} of the last predict; needs set_profiling(True) beforehand."""
        names = (ctypes.c_char_p * 32)()
        ms = (ctypes.c_float * 32)()
        cnt = (ctypes.c_int * 32)()
        n = self._lib.mrcnn_engine_kernel_times(self._engine, 32, names, ms, cnt)
        if n < 0:
            _native.check(n, "kernel_times")
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def step_table(self):
        """[(label, kernel family, ms, flops)] per launch of the last profiled predict."""
        out, i = [], 0
        while True:
            label, kind = ctypes.c_char_p(), ctypes.c_char_p()
            ms, fl = ctypes.c_float(), ctypes.c_double()
            if self._lib.mrcnn_engine_step_info(self._engine, i, ctypes.byref(label), ctypes.byref(kind), ctypes.byref(ms),
                                                ctypes.byref(fl)) != 0:
                break
            out.append((label.value.decode(), kind.value.decode(), float(ms.value), float(fl.value)))
            i += 1
        return out

    def set_profiling(self, on=True):
        """False/0 off; True/1 CUDA events around every launch; 2 events at kernel-family boundaries only
        (cheap; kernel_times() then averages the last <= 8 predicts)."""
        _native.check(self._lib.mrcnn_engine_set_profiling(self._engine, int(on)), "set_profiling")

    def flops_per_predict(self):
        return float(self._lib.mrcnn_engine_flops(self._engine))

    def detect(self, images, verbose=0):
        """Runs the detection pipeline (reference: mrcnn/model.py:2623-2704).
        images: list of BATCH_SIZE uint8 [H,W,3] arrays. Returns one dict per image with
        rois [N,4] int32, class_ids [N] int32, scores [N] float32, masks [H,W,N] bool."""
        assert self.mode == "inference", "Create model in inference mode."
        assert len(images) == self.config.BATCH_SIZE, "len(images) must be equal to BATCH_SIZE"
        if verbose:
            log("Processing {} images".format(len(images)))
            for image in images:
                log("image", image)
        torch = utils._torch()
        with torch.cuda.stream(self._stream):
            molded, metas, windows = self._mold_inputs_device(images)
        return self._detect_device(molded, metas, windows, [im.shape for im in images])

    def detect_molded(self, molded_images, image_metas, verbose=0):
        """reference: mrcnn/model.py:2706-2762 — inputs already molded; the window is the whole image."""
        assert self.mode == "inference", "Create model in inference mode."
        assert len(molded_images) == self.config.BATCH_SIZE, "Number of images must be equal to BATCH_SIZE"
        torch = utils._torch()
        molded = np.ascontiguousarray(np.stack(molded_images), dtype=np.float32)
        image_shape = molded[0].shape
        windows = np.array([[0, 0, image_shape[0], image_shape[1]]] * len(molded_images))
        return self._detect_device(torch.from_numpy(molded).pin_memory(), np.asarray(image_metas), windows,
                                   [image_shape] * len(molded_images))

    def unmold_detections(self, detections, mrcnn_mask, original_image_shape, image_shape, window):
        """reference: mrcnn/model.py:2558-2621, for one image, computed on the GPU (this model's device and stream)."""
        torch = utils._torch()
        lib = self._lib
        D = detections.shape[0]
        dev = "cuda:%d" % self._device
        H0, W0 = int(original_image_shape[0]), int(original_image_shape[1])
        with torch.cuda.stream(self._stream):
            d_det = torch.from_numpy(np.ascontiguousarray(detections, dtype=np.float32)).to(dev)
            d_mask = torch.from_numpy(np.ascontiguousarray(mrcnn_mask, dtype=np.float32)).to(dev)
            d_win = torch.from_numpy(np.ascontiguousarray(window, dtype=np.int32).reshape(1, 4)).to(dev)
            rois = torch.empty((D, 4), dtype=torch.int32, device=dev)
            cls = torch.empty((D,), dtype=torch.int32, device=dev)
            sc = torch.empty((D,), dtype=torch.float32, device=dev)
            cnt = torch.empty((1,), dtype=torch.int32, device=dev)
            masks = torch.empty((H0, W0, D), dtype=torch.uint8, device=dev)
            ws = torch.empty((lib.mrcnn_unmold_workspace_bytes(1, D),), dtype=torch.uint8, device=dev)
            orig = (ctypes.c_int * 2)(H0, W0)
            img = (ctypes.c_int * 2)(int(image_shape[0]), int(image_shape[1]))
            _native.check(lib.mrcnn_unmold_detections(_native.ptr(d_det), _native.ptr(d_mask), 1, D, mrcnn_mask.shape[1],
                                                      mrcnn_mask.shape[2], mrcnn_mask.shape[3], orig, img, _native.ptr(d_win),
                                                      _native.ptr(rois), _native.ptr(cls), _native.ptr(sc), _native.ptr(cnt),
                                                      _native.ptr(masks), _native.ptr(ws), ws.numel(),
                                                      self._stream.cuda_stream), "unmold_detections")
            n = int(cnt.item())
            full = masks.cpu().numpy()[:, :, :n].view(np.bool_) if n > 0 else np.empty((H0, W0, 0))
            return rois.cpu().numpy()[:n], cls.cpu().numpy()[:n], sc.cpu().numpy()[:n], full
