"""Analyzer: detection post-processing of the reference (mrcnn/analyze.py:580-2173), detect side only.

Mirrors the reference's `Analyzer` surface for the path that follows `model.detect()` on both CLI commands —
`predict()` :833-904, `extract_det_masks()` :1162-1423, `make_json_results()` :1866-1942,
`write_json_results()` :1945-1957, `merge_masks()` :2142, `extract_mask_connected_components()` :2148,
`are_mask_connected()` :2154 — with the same attribute names, option flags and output structures.

What runs where. The reference holds every detection as a full-frame numpy array and runs skimage labelling three
times per mask pair plus sklearn's jaccard_score on the flattened frames. Here the masks stay in HBM as bit-planes
(csrc/analyze.cu through the C ABI: mrcnn_masks_pack, mrcnn_planes_label, mrcnn_labels_select,
mrcnn_planes_pair_stats, mrcnn_planes_union, mrcnn_planes_area_bbox, mrcnn_planes_pixels, mrcnn_planes_unpack);
the host keeps only the tiny graph logic (score filter and ordering, the merge graph's DFS components, networkx
maximal cliques, selection), written with the same numpy / networkx calls as the reference so ties and scalar
types behave identically. `predict_maps()` is the device-resident extension: it runs `model.detect_maps(...,
device_only=True)` and post-processes the whole batch without ever copying the [H,W,100] masks to the host.

Not provided (SURVEY.md §8 marks them out of scope): ground-truth handling / performance metrics, drawing, DS9
regions. The "vertexes" key of the JSON objects (skimage.measure.find_contours in the reference) is computed from each
object's pixel list by a host-only C++ restatement of the scikit-image 0.15 algorithm (csrc/host_contours.cpp; parity
unpinned: scikit-image is absent here).
There is no CPU fallback: every mask operation goes through libmrcnn_b200.so and needs a CUDA device.
"""
import ctypes
import gc
import itertools
import json
import logging
import os
import time

import numpy as np

from . import _native, utils

logger = logging.getLogger(__name__)

# labels that extract_det_masks never splits into connected components (analyze.py:1223)
_UNSPLIT_LABELS = ("galaxy_C2", "galaxy_C3", "galaxy", "extended-multisland")


class Graph:
    """Undirected graph with DFS connected components (reference: mrcnn/graph.py). Components come out in order
    of their smallest vertex, members in pre-order of a depth-first walk that follows edges in insertion order."""

    def __init__(self, V):
        self.V = V
        self.adj = [[] for _ in range(V)]

    def addEdge(self, v, w):
        self.adj[v].append(w)
        self.adj[w].append(v)

    def connectedComponents(self):
        visited = [False] * self.V
        components = []
        for start in range(self.V):
            if visited[start]:
                continue
            members, stack = [], [iter((start,))]
            while stack:                       # explicit stack: same visiting order as the recursive walk
                for v in stack[-1]:
                    if not visited[v]:
                        visited[v] = True
                        members.append(v)
                        stack.append(iter(self.adj[v]))
                        break
                else:
                    stack.pop()
            components.append(members)
        return components


class _Frame:
    """One image's detections: device masks [H,W,depth] uint8 (address), host class ids / scores."""

    def __init__(self, masks_ptr, depth, n, class_ids, scores, keepalive=None):
        self.masks_ptr, self.depth, self.n = masks_ptr, depth, n
        self.class_ids, self.scores = class_ids, scores
        self.keepalive = keepalive


class _FrameResult:
    def __init__(self):
        self.masks_final, self.class_ids_final, self.class_names_final = [], [], []
        self.scores_final, self.bboxes, self.captions = [], [], []
        self.pixels = []          # per final object: int32 [npix,2] (y,x) with the image origin already added
        self.vertexes = None      # per final object: list of contours (filled by the batched path, else by build_json_results)


class MaskPlaneOps:
    """Thin host wrapper of the bit-plane entry points of the C ABI (device memory through torch)."""

    def __init__(self, device=0, stream=None):
        self.torch = utils._torch()
        self.lib = _native.lib()
        self.device = self.torch.device("cuda:%d" % int(device))
        self.stream = stream

    def _st(self):
        return ctypes.c_void_p(self.stream.cuda_stream if self.stream is not None
                               else self.torch.cuda.current_stream(self.device).cuda_stream)

    def _ctx(self):
        return self.torch.cuda.stream(self.stream) if self.stream is not None else self.torch.cuda.device(self.device)

    def _call(self, fn, *args, what=""):
        """Every native launch happens with this object's device current (the C entry points launch on the stream they
        are given and never switch devices themselves): a model on cuda:1 works while cuda:0 is the process default."""
        with self.torch.cuda.device(self.device):
            _native.check(fn(*args), what)

    def words(self, H, W):
        return int(self.lib.mrcnn_plane_words(H, W))

    def to_dev(self, array, dtype):
        t = self.torch.from_numpy(np.ascontiguousarray(array, dtype=dtype))
        with self._ctx():
            return t.to(self.device, non_blocking=False)

    def empty(self, shape, dtype):
        with self._ctx():
            return self.torch.empty(shape, dtype=dtype, device=self.device)

    def pack(self, masks_ptr, n_images, H, W, depth, plane_of, n_planes):
        """masks [n_images,H,W,depth] uint8 at device address masks_ptr; plane_of host int32 [n_images*depth]."""
        planes = self.empty((max(n_planes, 1), self.words(H, W)), self.torch.int32)
        d_map = self.to_dev(plane_of, np.int32)
        self._call(self.lib.mrcnn_masks_pack, ctypes.c_void_p(masks_ptr), n_images, H, W, depth, _native.ptr(d_map),
                                                _native.ptr(planes), self._st(), what="masks_pack")
        return planes[:n_planes]

    def area_bbox(self, planes, H, W):
        n = int(planes.shape[0])
        area = self.empty((max(n, 1),), self.torch.int32)
        bbox = self.empty((max(n, 1), 4), self.torch.int32)
        self._call(self.lib.mrcnn_planes_area_bbox, _native.ptr(planes), n, H, W, _native.ptr(area), _native.ptr(bbox),
                                                      self._st(), what="planes_area_bbox")
        return area[:n], bbox[:n]

    def pair_stats(self, planes, H, W, pairs, bbox=None):
        """pairs host int32 [P,2] -> device (inter [P], touch [P]); bbox: optional device [n,4] from area_bbox."""
        P = int(pairs.shape[0])
        inter = self.empty((max(P, 1),), self.torch.int32)
        touch = self.empty((max(P, 1),), self.torch.int32)
        if P:
            d_pairs = self.to_dev(pairs, np.int32)
            self._call(self.lib.mrcnn_planes_pair_stats, _native.ptr(planes), H, W, _native.ptr(d_pairs), P, _native.ptr(bbox),
                                                           _native.ptr(inter), _native.ptr(touch), self._st(), what="planes_pair_stats")
        return inter[:P], touch[:P]

    def union(self, planes, H, W, groups):
        """groups: list of lists of plane indices, or (members int32 [M], offsets int32 [G+1]) -> [G, words] planes."""
        if isinstance(groups, tuple):
            members, offsets = groups
            G = len(offsets) - 1
        else:
            G = len(groups)
            lens = np.fromiter(map(len, groups), dtype=np.int64, count=G)
            members = np.fromiter(itertools.chain.from_iterable(groups), dtype=np.int32, count=int(lens.sum()))
            offsets = np.zeros(G + 1, dtype=np.int32)
            offsets[1:] = np.cumsum(lens)
        out = self.empty((max(G, 1), self.words(H, W)), self.torch.int32)
        if G:
            d_m, d_o = self.to_dev(members, np.int32), self.to_dev(offsets, np.int32)
            self._call(self.lib.mrcnn_planes_union, _native.ptr(planes), H, W, _native.ptr(d_m), _native.ptr(d_o), G,
                                                      _native.ptr(out), self._st(), what="planes_union")
        return out[:G]

    def label(self, planes, H, W):
        """-> (labels [n,H,W] int32 device, counts [n] int32 device)"""
        n = int(planes.shape[0])
        labels = self.empty((max(n, 1), H, W), self.torch.int32)
        counts = self.empty((max(n, 1),), self.torch.int32)
        if n:
            ws_bytes = int(self.lib.mrcnn_planes_label_workspace_bytes(n, H, W))
            ws = self.empty((ws_bytes,), self.torch.uint8)
            self._call(self.lib.mrcnn_planes_label, _native.ptr(planes), n, H, W, _native.ptr(labels), _native.ptr(counts),
                                                      _native.ptr(ws), ws_bytes, self._st(), what="planes_label")
        return labels[:n], counts[:n]

    def select(self, labels, H, W, src, comp):
        K = len(src)
        out = self.empty((max(K, 1), self.words(H, W)), self.torch.int32)
        if K:
            d_s, d_c = self.to_dev(src, np.int32), self.to_dev(comp, np.int32)
            self._call(self.lib.mrcnn_labels_select, _native.ptr(labels), H, W, _native.ptr(d_s), _native.ptr(d_c), K,
                                                       _native.ptr(out), self._st(), what="labels_select")
        return out[:K]

    def pixels(self, planes, H, W, areas, y0=0, x0=0):
        """areas: host int array [n]; -> host int32 [sum(areas),2] and the int64 offsets [n+1]."""
        n = int(planes.shape[0])
        offsets = np.zeros(n + 1, dtype=np.int64)
        offsets[1:] = np.cumsum(np.asarray(areas, dtype=np.int64))
        total = int(offsets[-1])
        out = self.empty((max(total, 1), 2), self.torch.int32)
        if n and total:
            d_o = self.to_dev(offsets[:-1], np.int64)
            self._call(self.lib.mrcnn_planes_pixels, _native.ptr(planes), n, H, W, _native.ptr(d_o), int(y0), int(x0),
                                                       _native.ptr(out), self._st(), what="planes_pixels")
        return self.host(out[:total]), offsets

    def unpack(self, planes, H, W):
        n = int(planes.shape[0])
        out = self.empty((max(n, 1), H, W), self.torch.uint8)
        if n:
            self._call(self.lib.mrcnn_planes_unpack, _native.ptr(planes), n, H, W, _native.ptr(out), self._st(), what="planes_unpack")
        return self.host(out[:n])

    def concat(self, a, b):
        with self._ctx():
            return self.torch.cat([a, b], dim=0)

    def gather(self, planes, index):
        with self._ctx():
            idx = self.torch.from_numpy(np.asarray(index, dtype=np.int64)).to(self.device)
            return planes.index_select(0, idx).contiguous()

    def host(self, t):
        with self._ctx():
            h = t.cpu()
        return h.numpy()


class Analyzer(object):
    """Post-processes the detector output of one image into the final source list (reference: analyze.py:580)."""

    def __init__(self, model, config, dataset=None):
        self.model = model
        self.config = config
        self.n_classes = dataset.nclasses if dataset else self.config.NUM_CLASSES
        self.dataset = dataset
        # data
        self.image = None
        self.image_header = None
        self.image_id = -1
        self.image_xmin = 0
        self.image_ymin = 0
        # raw model output
        self.class_names = None
        self.masks = None
        self.boxes = None
        self.class_ids = None
        self.scores = None
        self.nobjects = 0
        # processed detections
        self.masks_final = []
        self.class_ids_final = []
        self.class_names_final = []
        self.scores_final = []
        self.bboxes = []
        self.captions = []
        self.split_masks = False
        self.merge_overlapped_masks = True
        self.select_best_overlapped_masks = True
        self.split_source_sidelobe = True
        self.merge_overlap_iou_thr = 0.3
        self.results = {}
        self.obj_name_tag = ""
        # thresholds
        self.score_thr = 0.7
        self.iou_thr = 0.6
        # outputs (drawing and DS9 regions are outside the rebuilt path: both default to off here)
        self.outfile = ""
        self.outfile_json = ""
        self.outfile_ds9 = ""
        self.draw = False
        self.write_to_json = True
        self.write_to_ds9 = False
        self._final_pixels = None
        self._ops = None
        # predict_maps / predict_maps_stream only: False returns each object's "pixels" as an int32 [npix,2] array
        # (same JSON through write_json_results / NumpyEncoder, far fewer Python objects per batch)
        self.pixels_as_lists = False
        # the "vertexes" key (contours of every catalogued object, mrcnn/analyze.py:1908-1927); False leaves it empty
        self.compute_vertexes = True

    # -- plumbing -------------------------------------------------------------------------------
    def _plane_ops(self):
        if self._ops is None:
            device = getattr(self.model, "_device", 0) if self.model is not None else 0
            stream = getattr(self.model, "_stream", None) if self.model is not None else None
            self._ops = MaskPlaneOps(device, stream)
        return self._ops

    def _options(self):
        return dict(score_thr=self.score_thr, split_masks=self.split_masks, merge_overlapped_masks=self.merge_overlapped_masks,
                    select_best_overlapped_masks=self.select_best_overlapped_masks,
                    split_source_sidelobe=self.split_source_sidelobe, merge_overlap_iou_thr=self.merge_overlap_iou_thr)

    # -- reference API ----------------------------------------------------------------------------
    def predict(self, image, image_id='', bboxes_gt=[], header=None, xmin=0, ymin=0):
        """reference: analyze.py:833-904 (drawing and DS9 output excluded)."""
        if image is None:
            logger.error("No input image given!")
            return -1
        if self.draw or self.write_to_ds9:
            raise NotImplementedError("drawing / DS9 regions are outside the rebuilt detect path (DESIGN.md §6): "
                                      "set analyzer.draw = False and analyzer.write_to_ds9 = False")
        self.image = image
        self.image_xmin = xmin
        self.image_ymin = ymin
        if image_id:
            self.image_id = image_id
        if header:
            self.image_header = header
        r = self.model.detect([self.image], verbose=0)[0]
        self.class_names = self.config.CLASS_NAMES
        self.masks = r['masks']
        self.boxes = r['rois']
        self.class_ids = r['class_ids']
        self.scores = r['scores']
        self.nobjects = self.masks.shape[-1]
        if self.nobjects > 0:
            self.extract_det_masks()
        else:
            logger.warning("No detected object found for image %s ..." % self.image_id)
            return 0
        self.bboxes_gt = bboxes_gt
        self.make_json_results()
        if self.write_to_json:
            self.write_json_results(self.outfile_json if self.outfile_json != "" else 'out_' + str(self.image_id) + '.json')
        return 0

    def extract_det_masks(self):
        """reference: analyze.py:1162-1423, from self.masks [H,W,N] / self.boxes / self.class_ids / self.scores."""
        masks = np.asarray(self.masks)
        H, W, depth = masks.shape
        N = int(np.asarray(self.boxes).shape[0])
        if depth == 0 or N == 0:                         # nothing detected: empty result lists, as in the reference
            self._publish(_FrameResult())
            return
        ops = self._plane_ops()
        d_masks = ops.to_dev(masks.view(np.uint8) if masks.dtype == np.bool_ else (masks != 0).view(np.uint8), np.uint8)
        frame = _Frame(d_masks.data_ptr(), depth, N, self.class_ids, self.scores, keepalive=d_masks)
        res = analyze_frames(ops, [frame], H, W, self.class_names, origins=[(self.image_ymin, self.image_xmin)],
                             **self._options())[0]
        self._publish(res)

    def _publish(self, res):
        self.masks_final = res.masks_final
        self.class_ids_final = res.class_ids_final
        self.class_names_final = res.class_names_final
        self.scores_final = res.scores_final
        self.bboxes = res.bboxes
        self.captions = res.captions
        self._final_pixels = res.pixels

    def make_json_results(self):
        """reference: analyze.py:1866-1942."""
        shape = self.image.shape
        self.results = build_json_results(self.image_id, self.obj_name_tag, self.class_names, shape[0], shape[1],
                                          self.image_xmin, self.image_ymin, self.masks_final, self.class_ids_final,
                                          self.scores_final, self.bboxes, self._final_pixels,
                                          compute_vertexes=self.compute_vertexes)

    def write_json_results(self, outfile):
        """reference: analyze.py:1945-1957 (numpy scalars are written as plain numbers)."""
        if not self.results:
            logger.warning("Result obj dictionary is empty, nothing to be written...")
            return
        with open(outfile, 'w') as fp:
            json.dump(self.results, fp, indent=2, sort_keys=True, cls=NumpyEncoder)

    def merge_masks(self, mask1, mask2):
        """reference: analyze.py:2142-2146 — union of two masks (computed on the device)."""
        ops = self._plane_ops()
        H, W = np.asarray(mask1).shape
        planes = _planes_from_host(ops, [mask1, mask2])
        out = ops.unpack(ops.union(planes, H, W, [[0, 1]]), H, W)[0]
        dt = np.result_type(np.asarray(mask1).dtype, np.asarray(mask2).dtype)
        return out.view(np.bool_) if dt == np.bool_ else out.astype(dt)

    def extract_mask_connected_components(self, mask):
        """reference: analyze.py:2148-2151 — (labels [H,W], ncomponents), 4-connectivity, raster-order numbering."""
        ops = self._plane_ops()
        H, W = np.asarray(mask).shape
        labels, counts = ops.label(_planes_from_host(ops, [mask]), H, W)
        return ops.host(labels)[0].astype(np.int64), int(ops.host(counts)[0])

    def are_mask_connected(self, mask1, mask2):
        """reference: analyze.py:2154-2173."""
        ops = self._plane_ops()
        H, W = np.asarray(mask1).shape
        _, touch = ops.pair_stats(_planes_from_host(ops, [mask1, mask2]), H, W, np.array([[0, 1]], dtype=np.int32))
        return bool(ops.host(touch)[0])

    # -- device-resident extension ---------------------------------------------------------------------
    def predict_maps(self, maps, image_ids=None, origins=None, zscale_contrasts=(0.25, 0.25, 0.25)):
        """maps [BATCH_SIZE,H,W] float32 -> one results dict (make_json_results layout) per image. The detector's
        [B,H,W,100] masks never leave the GPU: only class ids, scores, boxes and the final pixel lists do.
        "pixels" is an int32 [npix,2] array per object unless self.pixels_as_lists is set."""
        return self._finish_maps(self.model.detect_maps(maps, zscale_contrasts, masks_on_device=True), image_ids, origins, None)

    def predict_maps_stream(self, batches, zscale_contrasts=(0.25, 0.25, 0.25)):
        """Generator over an iterable of maps batches ([BATCH_SIZE,H,W] float32 each, or (maps, image_ids, origins[,
        name_tags]) tuples; name_tags = one obj_name_tag per image): yields predict_maps() results batch by batch, with the detector already working on batch k+1 while
        the post-processing of batch k (its few small kernels, on a stream of their own, and the host graph logic)
        runs. Two result slots exist in the engine, so at most two batches are ever in flight."""
        pending = None
        for item in batches:
            maps, image_ids, origins, name_tags = (tuple(item) + (None,))[:4] if isinstance(item, tuple) else (item, None, None, None)
            handle = (self.model.detect_maps(maps, zscale_contrasts, masks_on_device=True), image_ids, origins, name_tags)
            if pending is not None:
                yield self._finish_maps(*pending)
            pending = handle
        if pending is not None:
            yield self._finish_maps(*pending)

    def _side_stream_ops(self):
        """Plane ops on a stream of their own, so they do not queue behind the next batch's detect kernels."""
        if getattr(self, "_side_ops", None) is None:
            torch = utils._torch()
            device = getattr(self.model, "_device", 0)
            with torch.cuda.device(int(device)):
                self._side_ops = MaskPlaneOps(device, torch.cuda.Stream())
        return self._side_ops

    def _finish_maps(self, handle, image_ids, origins, name_tags=None):
        # The batch builds ~10^4 short-lived lists / scalars; with a large heap (torch, networkx) every generation-2
        # pass of the cyclic collector they trigger costs tens of ms. Nothing here creates reference cycles, so the
        # collector is paused for the duration (plain reference counting still frees everything).
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._finish_maps_impl(handle, image_ids, origins, name_tags)
        finally:
            if was_enabled:
                gc.enable()

    def _finish_maps_impl(self, handle, image_ids, origins, name_tags):
        c = self.config
        B = c.BATCH_SIZE
        H, W = handle.frame_hw
        tw = time.perf_counter()
        r = handle.result()                  # waits for this batch's small copies only (=> its masks are complete)
        if getattr(self, "_timings", None) is not None:
            self._timings["wait for detect"] = self._timings.get("wait for detect", 0.0) + time.perf_counter() - tw
        D, counts = r["depth"], r["counts"]
        frames = [_Frame(r["masks_ptr"] + b * H * W * D, D, int(counts[b]), r["class_ids"][b, :counts[b]],
                         r["scores"][b, :counts[b]]) for b in range(B)]
        origins = origins if origins is not None else [(0, 0)] * B
        self.class_names = c.CLASS_NAMES
        timings = getattr(self, "_timings", None)            # development aid (tools/catalog_bench.py)
        t0 = time.perf_counter()
        vmode = None if not self.compute_vertexes else ("lists" if self.pixels_as_lists else "arrays")
        if _array_path_applies(frames, self.split_masks):
            opts = self._options()
            batch = _analyze_batch_arrays(self._side_stream_ops(), frames, H, W, self.class_names, origins, False,
                                          timings=timings, vertexes=vmode, **opts)
            t1 = time.perf_counter()
            out = build_json_results_batch(batch, image_ids, name_tags if name_tags is not None else self.obj_name_tag,
                                           self.class_names, H, W, origins, self.pixels_as_lists)
            if timings is not None:
                timings["analyze_frames total"] = timings.get("analyze_frames total", 0.0) + t1 - t0
                timings["host: catalogue dicts"] = timings.get("host: catalogue dicts", 0.0) + time.perf_counter() - t1
            return out
        results = analyze_frames(self._side_stream_ops(), frames, H, W, self.class_names, origins=origins, want_masks=False,
                                 timings=timings, vertexes=vmode, **self._options())
        t1 = time.perf_counter()
        out = []
        for b, res in enumerate(results):
            image_id = image_ids[b] if image_ids is not None else b
            tag = name_tags[b] if name_tags is not None else self.obj_name_tag
            out.append(build_json_results(image_id, tag, self.class_names, H, W, origins[b][1], origins[b][0],
                                          res.masks_final, res.class_ids_final, res.scores_final, res.bboxes, res.pixels,
                                          pixels_as_lists=self.pixels_as_lists, compute_vertexes=self.compute_vertexes,
                                          vertexes=res.vertexes))
        if timings is not None:
            timings["analyze_frames total"] = timings.get("analyze_frames total", 0.0) + t1 - t0
            timings["host: catalogue dicts"] = timings.get("host: catalogue dicts", 0.0) + time.perf_counter() - t1
        return out


class ModelTester(object):
    """Ground-truth evaluation of the reference (mrcnn/analyze.py:65-575: mAP, confusion matrices, purity / completeness).
    Outside the rebuilt path (out of scope); the name is kept so that `from mrcnn.analyze import ModelTester`
    resolves.  The pieces it is built from are available: mrcnn.utils.compute_ap / compute_matches / compute_recall
    (pinned to the reference) and Analyzer for the detections."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("ModelTester (ground-truth metrics of the reference) is not provided by the B200 build; "
                                  "use mrcnn.utils.compute_ap / compute_matches on Analyzer results (DESIGN.md §6)")


class NumpyEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return json.JSONEncoder.default(self, obj)


def _planes_from_host(ops, masks):
    stack = np.stack([(np.asarray(m) != 0) for m in masks], axis=-1).view(np.uint8)      # [H,W,n]
    H, W, n = stack.shape
    d = ops.to_dev(stack, np.uint8)
    return ops.pack(d.data_ptr(), 1, H, W, n, np.arange(n, dtype=np.int32), n)


def contours_of_flat_pixels(flat, offsets, as_lists=True):
    """`vertexes` of objects whose pixel lists are slices flat[offsets[o]:offsets[o+1]] of ONE int32 [total,2] (y,x)
    array (image coordinates, origin included): per object the contours of find_contours(zero-padded mask, 0.5), each
    a list of [x, y] floats (as_lists) or a float64 [n,2] array view — reference mrcnn/analyze.py:1908-1927.  Host-only
    multi-threaded C++ behind the C ABI (csrc/host_contours.cpp: the scikit-image 0.15 algorithm restated;
    scikit-image itself is absent here, so this key is parity-UNPINNED)."""
    n = len(offsets) - 1
    if n <= 0:
        return []
    flat = np.ascontiguousarray(flat, dtype=np.int32).reshape(-1, 2)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lib = _native.lib()
    nv, nc = ctypes.c_int64(0), ctypes.c_int64(0)
    _native.check(lib.mrcnn_host_contours(flat.ctypes.data, offsets.ctypes.data, n, ctypes.byref(nv), ctypes.byref(nc)), "host_contours")
    verts = np.empty((nv.value, 2), dtype=np.float64)
    c_off = np.empty(nc.value + 1, dtype=np.int64)
    o_off = np.empty(n + 1, dtype=np.int64)
    _native.check(lib.mrcnn_host_contours_fetch(verts.ctypes.data, c_off.ctypes.data, o_off.ctypes.data), "host_contours_fetch")
    c_off, o_off = c_off.tolist(), o_off.tolist()
    if as_lists:
        return [[verts[c_off[k]:c_off[k + 1]].tolist() for k in range(o_off[o], o_off[o + 1])] for o in range(n)]
    return [[verts[c_off[k]:c_off[k + 1]] for k in range(o_off[o], o_off[o + 1])] for o in range(n)]


def contours_of_pixel_lists(pixel_lists, as_lists=True):
    """same for a list of separate pixel lists / arrays"""
    n = len(pixel_lists)
    if n == 0:
        return []
    arrays = [p if (isinstance(p, np.ndarray) and p.dtype == np.int32 and p.ndim == 2) else
              np.asarray(p, dtype=np.int32).reshape(-1, 2) for p in pixel_lists]
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(a) for a in arrays])
    flat = np.concatenate(arrays) if offsets[-1] else np.zeros((0, 2), dtype=np.int32)
    return contours_of_flat_pixels(flat, offsets, as_lists)


def build_json_results(image_id, obj_name_tag, class_names, ny, nx, xmin, ymin, masks_final, class_ids_final, scores_final,
                       bboxes, pixels, pixels_as_lists=True, compute_vertexes=True, vertexes=None):
    """reference: analyze.py:1866-1942. `pixels`: per object int32 [npix,2] (y,x), image origin already added
    (np.argwhere(mask==1) computed on the device). pixels_as_lists=False keeps each object's "pixels" as that int32
    array instead of a list of [y, x] lists (NumpyEncoder writes the same JSON): a batch of 64 images otherwise
    allocates ~10^5 small lists, which costs more in Python's garbage collector than the whole GPU step."""
    n = len(class_ids_final)
    results = {"image_id": image_id, "objs": []}
    if n == 0:
        return results
    if vertexes is None and compute_vertexes:       # (the batched path computes them for all frames in one call)
        vertexes = contours_of_pixel_lists([pixels[i] for i in range(n)], as_lists=pixels_as_lists)
    # per-frame conversions in one go (a 64-frame batch holds ~2 500 objects: per-object numpy scalar handling was the
    # largest host item of the catalogue path)
    bb = np.asarray(bboxes, dtype=np.int64).reshape(n, 4)
    edge = ((bb[:, [1, 3]] <= 0) | (bb[:, [1, 3]] >= nx - 1)).any(axis=1) | ((bb[:, [0, 2]] <= 0) | (bb[:, [0, 2]] >= ny - 1)).any(axis=1)
    edge = edge.tolist()
    bb = (bb + np.array([ymin, xmin, ymin, xmin], dtype=np.int64)).tolist()
    cids = [int(c) for c in class_ids_final]
    objs = results["objs"]
    for i in range(n):
        y1, x1, y2, x2 = bb[i]
        class_id = cids[i]
        objs.append({
            "name": "S%d_%s" % (i + 1, obj_name_tag),
            "x1": x1, "x2": x2, "y1": y1, "y2": y2,
            "class_id": class_id, "class_name": class_names[class_id], "score": scores_final[i],
            "pixels": pixels[i].tolist() if pixels_as_lists else pixels[i],
            "vertexes": vertexes[i] if vertexes is not None else [], "edge": edge[i],
        })
    return results


def build_json_results_batch(batch, image_ids, name_tags, class_names, ny, nx, origins, pixels_as_lists=False):
    """build_json_results for every frame of a _BatchArrays in one pass (reference: analyze.py:1866-1942 per image): the
    box / edge arithmetic and the scalar conversions run once over the batch. name_tags: one tag or one per frame."""
    F = len(batch.frame_off) - 1
    ids = image_ids if image_ids is not None else range(F)
    out = [{"image_id": ids[f], "objs": []} for f in range(F)]
    n = int(batch.frame_off[-1])
    if not batch.published or n == 0:
        return out
    bb = batch.bbox.astype(np.int64)
    edge = (((bb[:, [1, 3]] <= 0) | (bb[:, [1, 3]] >= nx - 1)).any(axis=1) | ((bb[:, [0, 2]] <= 0) | (bb[:, [0, 2]] >= ny - 1)).any(axis=1)).tolist()
    per_frame = np.diff(batch.frame_off)
    org = np.asarray([(o[0], o[1], o[0], o[1]) for o in origins], dtype=np.int64).reshape(F, 4)
    bb = (bb + np.repeat(org, per_frame, axis=0)).tolist()
    cids = batch.cls.tolist()
    scores = list(batch.score)                      # numpy float32 scalars, as the reference's results carry
    po = batch.px_off.tolist()
    px = batch.px
    vx = batch.vertexes
    fo = batch.frame_off.tolist()
    for f in range(F):
        a, b = fo[f], fo[f + 1]
        if a == b:
            continue
        tag = name_tags if isinstance(name_tags, str) else name_tags[f]
        objs = out[f]["objs"]
        for k in range(a, b):
            y1, x1, y2, x2 = bb[k]
            class_id = cids[k]
            pix = px[po[k]:po[k + 1]]
            objs.append({
                "name": "S%d_%s" % (k - a + 1, tag),
                "x1": x1, "x2": x2, "y1": y1, "y2": y2,
                "class_id": class_id, "class_name": class_names[class_id], "score": scores[k],
                "pixels": pix.tolist() if pixels_as_lists else pix,
                "vertexes": vx[k] if vx is not None else [], "edge": edge[k],
            })
    return out


_TRIU = {}


def _all_pairs(counts):
    """(i<j) pairs inside each frame, in the reference's loop order -> (pairs [P,2] global indices, per-frame slices)."""
    chunks, slices, base, pos = [], [], 0, 0
    for n in counts:
        if n > 1:
            if n not in _TRIU:
                i, j = np.triu_indices(n, k=1)
                _TRIU[n] = np.stack([i, j], axis=1).astype(np.int32)
            chunks.append(_TRIU[n] + np.int32(base))
        npairs = n * (n - 1) // 2
        slices.append((pos, pos + npairs, base))
        pos += npairs
        base += n
    pairs = np.concatenate(chunks) if chunks else np.zeros((0, 2), dtype=np.int32)
    return pairs, slices


def _probe_f32_average():
    score_avg = 0
    score_avg += np.float32(0.8125)
    score_avg *= 1. / 1
    return type(score_avg) is np.float32 and score_avg == np.float32(0.8125)


# True when the reference's `score_avg = 0; score_avg += s; score_avg *= 1./1` leaves a float32 scalar unchanged
# (numpy >= 2 scalar promotion); otherwise the expression is always evaluated literally
_F32_AVG_IS_IDENTITY = _probe_f32_average()


def _probe_scalar_compare():
    x = np.float32(0.7)
    return bool(x < 0.7) == bool((np.array([x]) < 0.7)[0]) and bool(x < 0.7000000001) == bool((np.array([x]) < 0.7000000001)[0])


# True when `np.float32 scalar < python float` and `float32 array < python float` compare the same way (numpy >= 2:
# both in float32); numpy 1.x compares the scalar in float64, and then the reference's scalar loop is kept
_SCALAR_LT_IS_ARRAY_LT = _probe_scalar_compare()


def _iou(inter, area_a, area_b):
    """sklearn jaccard_score(average='binary'): tp / (tp + fp + fn) in float64, 0.0 for an empty union."""
    union = area_a.astype(np.int64) + area_b.astype(np.int64) - inter.astype(np.int64)
    out = np.zeros(inter.shape, dtype=np.float64)
    np.divide(inter.astype(np.float64), union.astype(np.float64), out=out, where=union > 0)
    return out


# MRCNN_B200_ANALYZE_GENERIC=1 forces the per-frame Python walk below for every call (the array path is the default
# wherever it applies; both are replayed against the reference goldens by the CPU suite)
_FORCE_GENERIC = os.environ.get("MRCNN_B200_ANALYZE_GENERIC", "0") == "1"


class _BatchArrays:
    """Result of the array path for a whole batch: the final objects of all frames as flat arrays, frame-major.
    frame_off [F+1] object ranges; cls / score / bbox per object; px flat int32 [total,2] pixel lists (origin added) with
    px_off [n+1]; vertexes per object (or None); masks / mask_is_int only with want_masks."""

    def __init__(self, F):
        self.frame_off = np.zeros(F + 1, dtype=np.int64)
        self.cls = np.zeros(0, dtype=np.int32)
        self.score = np.zeros(0, dtype=np.float32)
        self.bbox = np.zeros((0, 4), dtype=np.int32)
        self.px = np.zeros((0, 2), dtype=np.int32)
        self.px_off = np.zeros(1, dtype=np.int64)
        self.vertexes = None
        self.masks = None
        self.published = True          # False: select_best_overlapped_masks off -> nothing is published (analyze.py:1324)

    def frame_results(self, class_names, want_captions=True):
        """-> one _FrameResult per frame (lists, as the reference's attributes)"""
        F = len(self.frame_off) - 1
        out = [_FrameResult() for _ in range(F)]
        if not self.published:
            return out
        fo = self.frame_off.tolist()
        po = self.px_off.tolist()
        for f in range(F):
            a, b = fo[f], fo[f + 1]
            if a == b:
                continue
            res = out[f]
            res.class_ids_final = list(self.cls[a:b])
            res.class_names_final = [class_names[c] for c in self.cls[a:b].tolist()]
            res.scores_final = list(self.score[a:b])
            res.bboxes = list(self.bbox[a:b])
            if want_captions:
                res.captions = ["{} {:.2f}".format(n, sc) for n, sc in zip(res.class_names_final, res.scores_final)]
            res.pixels = [self.px[po[k]:po[k + 1]] for k in range(a, b)]
            res.masks_final = [self.masks[k] for k in range(a, b)] if self.masks is not None else [None] * (b - a)
            if self.vertexes is not None:
                res.vertexes = self.vertexes[a:b]
        return out


def _array_path_applies(frames, split_masks=False):
    if _FORCE_GENERIC or not (_F32_AVG_IS_IDENTITY and _SCALAR_LT_IS_ARRAY_LT):
        return False
    return all(isinstance(fr.scores, np.ndarray) and fr.scores.dtype == np.float32 and isinstance(fr.class_ids, np.ndarray)
               and fr.class_ids.dtype.kind in "iu" for fr in frames)


def _host_pairs(counts):
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    c64 = counts.astype(np.int64)
    P = int((c64 * (c64 - 1) // 2).sum())
    pairs = np.empty((P, 2), dtype=np.int32)
    _native.check(_native.lib().mrcnn_host_all_pairs(len(counts), counts.ctypes.data, pairs.ctypes.data if P else None),
                  "host_all_pairs")
    return counts, pairs


def _pair_flags(stage, counts, key, score, area, inter, touch, use_iou, thr, n_total):
    P = len(inter)
    flags = np.empty(P, dtype=np.uint8)
    loses = np.zeros(n_total, dtype=np.uint8)
    tie = np.zeros(len(counts), dtype=np.uint8)
    key = np.ascontiguousarray(key, dtype=np.int32)
    area = np.ascontiguousarray(area, dtype=np.int32)
    inter = np.ascontiguousarray(inter, dtype=np.int32)
    touch = np.ascontiguousarray(touch, dtype=np.int32)
    score = None if score is None else np.ascontiguousarray(score, dtype=np.float32)
    _native.check(_native.lib().mrcnn_host_pair_flags(stage, len(counts), counts.ctypes.data, key.ctypes.data,
                                                      None if score is None else score.ctypes.data, area.ctypes.data,
                                                      inter.ctypes.data, touch.ctypes.data, 1 if use_iou else 0, float(thr),
                                                      flags.ctypes.data, loses.ctypes.data, tie.ctypes.data), "host_pair_flags")
    return flags, loses, tie


def _clique_selection(n, edges, scores):
    """analyze.py:1362-1395 for one frame: keep a mask iff it is the best-scoring member of every maximal clique it
    belongs to; cliques from networkx in its own order (the order decides between equal scores)."""
    import networkx as nx
    is_selected = [True] * n
    if not len(edges):
        return is_selected
    g_final = nx.Graph()
    g_final.add_edges_from(edges)                       # same insertion order as the reference's pair loop
    cliques = list(nx.find_cliques(g_final))
    clique_max_scores, clique_max_score_index = [], []
    for item in cliques:
        max_score, max_score_index = -1, -1
        for index in item:
            score = scores[index]
            if score > max_score:
                max_score, max_score_index = score, index
        clique_max_scores.append(max_score)
        clique_max_score_index.append(max_score_index)
    for q in sorted(range(len(cliques)), key=lambda k: clique_max_scores[k], reverse=True):
        for index in cliques[q]:
            if index != clique_max_score_index[q] and is_selected[index]:
                is_selected[index] = False
    return is_selected


def _analyze_batch_arrays(ops, frames, H, W, class_names, origins, want_masks, score_thr, merge_overlapped_masks,
                          select_best_overlapped_masks, split_source_sidelobe, merge_overlap_iou_thr, timings, vertexes,
                          split_masks=False):
    """extract_det_masks + pixel lists for a batch with every per-mask quantity held in ONE array over the batch: the
    per-frame work left in Python is one argsort (the reference's own call, so ties fall the same way); the pair tests,
    the merge components and the pair enumeration are host C++ (csrc/host_graph.cu), everything else is numpy over the
    batch. Same results as the per-frame walk of analyze_frames (replayed against the reference goldens by the CPU suite);
    frames in which two linked masks tie take the reference's networkx route."""
    import time

    def mark(stage):
        if timings is not None:
            ops.torch.cuda.synchronize()
            now = time.perf_counter()
            timings[stage] = timings.get(stage, 0.0) + now - mark.t
            mark.t = now
    mark.t = time.perf_counter()

    F = len(frames)
    depth = frames[0].depth
    out = _BatchArrays(F)
    # -- score filter and descending-score order (analyze.py:1181-1203)
    idx_chunks, cls_chunks, score_chunks, sel_count = [], [], [], []
    for f, fr in enumerate(frames):
        sc = fr.scores[:fr.n]
        picked = np.nonzero(~(sc < score_thr))[0]
        chosen = picked[np.argsort(sc[picked])[::-1]]
        idx_chunks.append(chosen + f * depth)
        cls_chunks.append(fr.class_ids[:fr.n][chosen])
        score_chunks.append(sc[chosen])
        sel_count.append(len(chosen))
    flat_idx = np.concatenate(idx_chunks)
    m = len(flat_idx)
    plane_of = np.full(F * depth, -1, dtype=np.int32)
    plane_of[flat_idx] = np.arange(m, dtype=np.int32)
    cls_arr = np.concatenate(cls_chunks)
    score_arr = np.concatenate(score_chunks)
    counts = np.asarray(sel_count, dtype=np.int32)
    mark("host: score filter + order")
    planes = ops.pack(frames[0].masks_ptr, F, H, W, depth, plane_of, m)
    mark("gpu: pack")

    # -- optional split into 4-connected components (analyze.py:1211-1255): a mask of a splittable class becomes one mask
    #    per component (none if it is empty), each an int64 array in the reference (np.where(labels == c + 1, [1], [0]))
    is_int = np.zeros(m, dtype=bool)
    if split_masks and m:
        unsplit = np.fromiter((name in _UNSPLIT_LABELS for name in class_names), dtype=bool, count=len(class_names))
        splittable = ~unsplit[cls_arr]
        labels, d_ncomp = ops.label(planes, H, W)
        ncomp = np.asarray(ops.host(d_ncomp)).astype(np.int64)
        mark("gpu: label components (+counts D2H)")
        k = np.where(splittable, ncomp, 1)
        src_all = np.repeat(np.arange(m, dtype=np.int64), k)           # det mask -> the selected mask it comes from
        first = np.cumsum(k) - k
        comp_all = np.arange(len(src_all), dtype=np.int64) - np.repeat(first, k) + 1
        from_split = splittable[src_all]
        parts = ops.select(labels, H, W, src_all[from_split].astype(np.int32), comp_all[from_split].astype(np.int32))
        if from_split.all():
            planes = parts
        else:
            origin_idx = np.empty(len(src_all), dtype=np.int64)         # row of cat([parts, planes]) for every det mask
            origin_idx[from_split] = np.arange(int(from_split.sum()))
            origin_idx[~from_split] = int(from_split.sum()) + src_all[~from_split]
            planes = ops.gather(ops.concat(parts, planes), origin_idx)
        frame_of_sel = np.repeat(np.arange(F), counts)
        counts = np.bincount(frame_of_sel, weights=k, minlength=F).astype(np.int32)
        cls_arr, score_arr, is_int = cls_arr[src_all], score_arr[src_all], from_split
        m = len(src_all)
        mark("host+gpu: component planes")

    # -- merge connected same-class masks above the IOU threshold (analyze.py:1258-1320)
    if merge_overlapped_masks and m:
        counts, pairs = _host_pairs(counts)
        mark("host: pair lists")
        d_area, d_bbox = ops.area_bbox(planes, H, W)
        d_inter, d_touch = ops.pair_stats(planes, H, W, pairs, d_bbox)
        area, inter, touch = ops.host(d_area), ops.host(d_inter), ops.host(d_touch)
        mark("gpu: merge pair stats (+pairs H2D, results D2H)")
        mergeable, _, _ = _pair_flags(0, counts, cls_arr, None, area, inter, touch, True, merge_overlap_iou_thr, m)
        members = np.empty(m, dtype=np.int32)
        offsets = np.empty(m + 1, dtype=np.int32)
        frame_comps = np.empty(F, dtype=np.int32)
        ncomp = ctypes.c_int32(0)
        _native.check(_native.lib().mrcnn_host_merge_components(
            F, counts.ctypes.data, pairs.ctypes.data if len(pairs) else None, mergeable.ctypes.data if len(pairs) else None,
            len(pairs), members.ctypes.data, offsets.ctypes.data, frame_comps.ctypes.data, ctypes.byref(ncomp)),
            "host_merge_components")
        G = ncomp.value
        offsets = offsets[:G + 1]
        if G < m:                                        # real merges exist
            sizes = np.diff(offsets)
            new_cls = cls_arr[members[offsets[1:] - 1]]      # class of the LAST member, as in the reference
            new_score = score_arr[members[offsets[:-1]]]     # singletons: (0 + s) * (1. / 1) is s
            for g in np.nonzero(sizes > 1)[0].tolist():      # the reference's scalar arithmetic, literally
                score_avg = 0
                for k in members[offsets[g]:offsets[g + 1]].tolist():
                    score_avg += score_arr[k]
                score_avg *= 1. / int(sizes[g])
                new_score[g] = score_avg
            if is_int.any():
                is_int = np.logical_or.reduceat(is_int[members], offsets[:-1])
            else:
                is_int = np.zeros(G, dtype=bool)
            cls_arr, score_arr, counts = new_cls, new_score, frame_comps
            mark("host: merge graph")
            planes = ops.union(planes, H, W, (members, offsets))
            mark("gpu: union")
        else:
            mark("host: merge graph")
    n_merged = len(cls_arr)
    if not select_best_overlapped_masks or not n_merged:
        out.published = bool(select_best_overlapped_masks)
        return out

    # -- best of overlapping objects through maximal cliques (analyze.py:1328-1395)
    counts, pairs = _host_pairs(counts)
    mark("host: pair lists")
    d_area, d_bbox = ops.area_bbox(planes, H, W)
    d_inter, d_touch = ops.pair_stats(planes, H, W, pairs, d_bbox)
    area, bbox, inter, touch = ops.host(d_area), ops.host(d_bbox), ops.host(d_inter), ops.host(d_touch)
    mark("gpu: select pair stats + bbox (+H2D/D2H)")
    spurious = np.fromiter((name == 'spurious' for name in class_names), dtype=np.int32, count=len(class_names))[cls_arr]
    linked, loses, tie_frame = _pair_flags(1, counts, spurious, score_arr, area, inter, touch, split_source_sidelobe,
                                           merge_overlap_iou_thr, n_merged)
    selected = loses == 0
    if tie_frame.any():
        starts = np.concatenate([[0], np.cumsum(counts)])
        c64 = counts.astype(np.int64)
        pstarts = np.concatenate([[0], np.cumsum(c64 * (c64 - 1) // 2)])
        for f in np.nonzero(tie_frame)[0].tolist():
            base, n_f = int(starts[f]), int(counts[f])
            lo, hi = int(pstarts[f]), int(pstarts[f + 1])
            edges = (pairs[lo:hi][linked[lo:hi] != 0] - base).tolist()
            selected[base:base + n_f] = _clique_selection(n_f, edges, list(score_arr[base:base + n_f]))
    bbox_ok = (bbox[:, 1] < bbox[:, 3]) & (bbox[:, 0] < bbox[:, 2])
    for k in np.nonzero(selected & ~bbox_ok)[0].tolist():
        bb = bbox[k]
        logger.warning("Invalid det bbox(%d,%d,%d,%d), skip it ..." % (bb[1], bb[3], bb[0], bb[2]))
    keep = np.nonzero(selected & bbox_ok)[0]
    owner = np.repeat(np.arange(F), counts)[keep]
    out.frame_off[1:] = np.cumsum(np.bincount(owner, minlength=F))
    out.cls, out.score, out.bbox = cls_arr[keep], score_arr[keep], bbox[keep]
    mark("host: cliques + selection")

    # -- final masks and pixel lists (make_json_results: np.argwhere(mask == 1), analyze.py:1903-1909)
    if len(keep):
        fin = planes if len(keep) == n_merged else ops.gather(planes, keep)
        if want_masks:
            dense = ops.unpack(fin, H, W)
            if is_int[keep].any():                                   # masks that went through the component split are int64
                out.masks = [dense[j].astype(np.int64) if flag else dense[j].view(np.bool_) for j, flag in enumerate(is_int[keep].tolist())]
            else:
                out.masks = dense.view(np.bool_)
        px, px_off = ops.pixels(fin, H, W, area[keep], 0, 0)          # one launch and one copy for the whole batch
        origins = origins if origins is not None else [(0, 0)] * F
        fo = out.frame_off.tolist()
        po = px_off.tolist()
        for f in range(F):
            oy, ox = int(origins[f][0]), int(origins[f][1])
            if (oy or ox) and fo[f] != fo[f + 1]:
                px[po[fo[f]]:po[fo[f + 1]]] += np.array([oy, ox], dtype=np.int32)
        out.px, out.px_off = px, px_off
        if vertexes is not None:                                         # contours of every final object, one threaded C++ call
            out.vertexes = contours_of_flat_pixels(px, px_off, as_lists=(vertexes == "lists"))
    mark("gpu+host: pixel lists / final masks")
    return out


def analyze_frames(ops, frames, H, W, class_names, origins=None, want_masks=True, score_thr=0.7, split_masks=False,
                   merge_overlapped_masks=True, select_best_overlapped_masks=True, split_source_sidelobe=True,
                   merge_overlap_iou_thr=0.3, timings=None, vertexes=None):
    """extract_det_masks (+ the pixel lists of make_json_results) for a list of frames that share one [H,W] size.
    All masks of one frame list must live in ONE device allocation laid out [n_frames,H,W,depth] when
    len(frames) > 1 (the engine's result slot), or be a single frame."""
    import networkx as nx
    import time

    if len(frames) and _array_path_applies(frames, split_masks):
        depth, base_ptr = frames[0].depth, frames[0].masks_ptr
        for f, fr in enumerate(frames):
            assert fr.depth == depth and fr.masks_ptr == base_ptr + f * H * W * depth, "frames must be one [F,H,W,depth] block"
        return _analyze_batch_arrays(ops, frames, H, W, class_names, origins, want_masks, score_thr, merge_overlapped_masks,
                                     select_best_overlapped_masks, split_source_sidelobe, merge_overlap_iou_thr, timings,
                                     vertexes, split_masks).frame_results(class_names)

    def mark(stage):          # development aid: cumulative wall time per stage (device drained at every mark)
        if timings is not None:
            ops.torch.cuda.synchronize()
            now = time.perf_counter()
            timings[stage] = timings.get(stage, 0.0) + now - mark.t
            mark.t = now
    mark.t = time.perf_counter()

    F = len(frames)
    depth = frames[0].depth
    base_ptr = frames[0].masks_ptr
    for f, fr in enumerate(frames):
        assert fr.depth == depth and fr.masks_ptr == base_ptr + f * H * W * depth, "frames must be one [F,H,W,depth] block"
    origins = origins if origins is not None else [(0, 0)] * F

    # -- score filter and descending-score order (analyze.py:1181-1203): host, the reference's own numpy calls
    plane_of = np.full(F * depth, -1, dtype=np.int32)
    sel_cls, sel_score, sel_count = [], [], []
    m = 0
    for f, fr in enumerate(frames):
        if _SCALAR_LT_IS_ARRAY_LT and isinstance(fr.scores, np.ndarray) and fr.scores.dtype == np.float32:
            picked = np.nonzero(~(fr.scores[:fr.n] < score_thr))[0].tolist()     # same comparisons, one numpy call
        else:
            picked = [i for i in range(fr.n) if not fr.scores[i] < score_thr]
        scores_sel = [fr.scores[i] for i in picked]
        order = np.argsort(scores_sel)[::-1]
        for index in order:
            plane_of[f * depth + picked[index]] = m
            sel_cls.append(fr.class_ids[picked[index]])
            sel_score.append(scores_sel[index])
            m += 1
        sel_count.append(len(picked))
    mark("host: score filter + order")
    planes = ops.pack(base_ptr, F, H, W, depth, plane_of, m)
    mark("gpu: pack")

    # -- optional split into 4-connected components (analyze.py:1211-1255)
    det_cls, det_score, det_int, det_count = sel_cls, sel_score, [False] * m, sel_count
    if split_masks and m:
        splittable = [class_names[c] not in _UNSPLIT_LABELS for c in sel_cls]
        labels, counts = ops.label(planes, H, W)
        ncomp = ops.host(counts)
        mark("gpu: label components (+counts D2H)")
        src, comp, keep_src, keep_dst = [], [], [], []
        det_cls, det_score, det_int, det_count = [], [], [], []
        pos = 0
        for f in range(F):
            n_f = 0
            for _ in range(sel_count[f]):
                k = 1 if not splittable[pos] else int(ncomp[pos])
                for c in range(k):
                    if splittable[pos]:
                        src.append(pos)
                        comp.append(c + 1)
                    else:
                        keep_src.append(pos)
                        keep_dst.append(len(det_cls))
                    det_cls.append(sel_cls[pos])
                    det_score.append(sel_score[pos])
                    det_int.append(splittable[pos])      # np.where(labels == c + 1, [1], [0]) is an int64 array
                n_f += k
                pos += 1
            det_count.append(n_f)
        parts = ops.select(labels, H, W, src, comp)
        if keep_src:
            is_keep = np.zeros(len(det_cls), dtype=bool)
            is_keep[keep_dst] = True
            origin_idx = np.empty(len(det_cls), dtype=np.int64)     # row of cat([parts, planes]) for every det mask
            origin_idx[~is_keep] = np.arange(len(src))
            origin_idx[is_keep] = len(src) + np.asarray(keep_src, dtype=np.int64)
            planes = ops.gather(ops.concat(parts, planes), origin_idx)
        else:
            planes = parts
        mark("host+gpu: component planes")

    # -- merge connected same-class masks above the IOU threshold (analyze.py:1258-1320)
    merged_cls, merged_score, merged_int, merged_count = det_cls, det_score, det_int, det_count
    if merge_overlapped_masks and len(det_cls):
        pairs, slices = _all_pairs(det_count)
        mark("host: pair lists")
        d_area, d_bbox = ops.area_bbox(planes, H, W)
        d_inter, d_touch = ops.pair_stats(planes, H, W, pairs, d_bbox)
        area, inter, touch = ops.host(d_area), ops.host(d_inter), ops.host(d_touch)
        mark("gpu: merge pair stats (+pairs H2D, results D2H)")
        cls_arr = np.asarray(det_cls)
        iou = _iou(inter, area[pairs[:, 0]], area[pairs[:, 1]])
        mergeable = (touch != 0) & (cls_arr[pairs[:, 0]] == cls_arr[pairs[:, 1]]) & (iou >= merge_overlap_iou_thr)
        groups, merged_cls, merged_score, merged_int, merged_count = [], [], [], [], []
        any_int = any(det_int)
        for f, (lo, hi, base) in enumerate(slices):
            edges = np.nonzero(mergeable[lo:hi])[0]
            if len(edges):
                g = Graph(det_count[f])
                for a, b in (pairs[lo + edges] - base).tolist():
                    g.addEdge(a, b)
                cc = g.connectedComponents()
            elif _F32_AVG_IS_IDENTITY and all(type(x) is np.float32 for x in det_score[base:base + det_count[f]]):
                # no edges: every mask is its own component and (0 + s) * (1. / 1) is s itself -> bulk copy
                n_f = det_count[f]
                groups.extend([k] for k in range(base, base + n_f))
                merged_cls.extend(det_cls[base:base + n_f])
                merged_score.extend(det_score[base:base + n_f])
                merged_int.extend(det_int[base:base + n_f] if any_int else [False] * n_f)
                merged_count.append(n_f)
                continue
            else:
                cc = [[v] for v in range(det_count[f])]        # what the DFS returns for a graph without edges
            for members in cc:
                if len(members) == 1 and _F32_AVG_IS_IDENTITY and type(det_score[base + members[0]]) is np.float32:
                    # (0 + s) * (1. / 1) is s itself, same scalar type, under this numpy's promotion rules
                    k = base + members[0]
                    groups.append([k])
                    merged_cls.append(det_cls[k])
                    merged_score.append(det_score[k])
                    merged_int.append(any_int and det_int[k])
                    continue
                score_avg = 0
                for index in members:
                    class_id = det_cls[base + index]
                    score_avg += det_score[base + index]
                score_avg *= 1. / len(members)
                groups.append([base + index for index in members])
                merged_cls.append(class_id)
                merged_score.append(score_avg)
                merged_int.append(any_int and any(det_int[base + index] for index in members))
            merged_count.append(len(cc))
        mark("host: merge graph")
        planes = ops.union(planes, H, W, groups)
        mark("gpu: union")

    results = [_FrameResult() for _ in range(F)]
    if not select_best_overlapped_masks or not len(merged_cls):     # analyze.py:1324: nothing is published otherwise
        return results

    # -- best of overlapping objects through maximal cliques (analyze.py:1328-1395)
    pairs, slices = _all_pairs(merged_count)
    mark("host: pair lists")
    d_area, d_bbox = ops.area_bbox(planes, H, W)
    d_inter, d_touch = ops.pair_stats(planes, H, W, pairs, d_bbox)
    area, bbox, inter, touch = ops.host(d_area), ops.host(d_bbox), ops.host(d_inter), ops.host(d_touch)
    mark("gpu: select pair stats + bbox (+H2D/D2H)")
    spurious = np.array([class_names[c] == 'spurious' for c in merged_cls], dtype=bool)
    linked = touch != 0
    if split_source_sidelobe:
        iou = _iou(inter, area[pairs[:, 0]], area[pairs[:, 1]])
        linked &= ~((spurious[pairs[:, 0]] != spurious[pairs[:, 1]]) & (iou < merge_overlap_iou_thr))
    final_planes, final_owner = [], []
    bbox_ok = ((bbox[:, 1] < bbox[:, 3]) & (bbox[:, 0] < bbox[:, 2])).tolist()
    # A mask survives iff it has the best score in every maximal clique it belongs to. Without score ties between
    # linked masks that is "no linked mask has a higher score" (any linked pair extends to a maximal clique that holds
    # both), which needs no clique enumeration; frames with a tie between linked masks take the reference's
    # networkx route below, where the winner depends on the order networkx lists a clique's members.
    score_arr = np.asarray(merged_score)
    li, lj = pairs[linked, 0], pairs[linked, 1]
    si, sj = score_arr[li], score_arr[lj]
    loses = np.zeros(len(merged_score), dtype=bool)
    loses[li[si < sj]] = True
    loses[lj[sj < si]] = True
    tie_frame = np.zeros(F, dtype=bool)
    if len(li):
        frame_of = np.repeat(np.arange(F), merged_count)
        tie_frame[frame_of[li[si == sj]]] = True
    for f, (lo, hi, base) in enumerate(slices):
        if not tie_frame[f]:
            is_selected = (~loses[base:base + merged_count[f]]).tolist()
            edges = ()
        else:
            is_selected = [True] * merged_count[f]
            edges = np.nonzero(linked[lo:hi])[0]
        if len(edges):
            g_final = nx.Graph()
            g_final.add_edges_from((pairs[lo + edges] - base).tolist())     # same insertion order as the pair loop
            cliques = list(nx.find_cliques(g_final))
            clique_max_scores, clique_max_score_index = [], []
            for item in cliques:
                max_score, max_score_index = -1, -1
                for index in item:
                    score = merged_score[base + index]
                    if score > max_score:
                        max_score, max_score_index = score, index
                clique_max_scores.append(max_score)
                clique_max_score_index.append(max_score_index)
            for q in sorted(range(len(cliques)), key=lambda k: clique_max_scores[k], reverse=True):
                for index in cliques[q]:
                    if index != clique_max_score_index[q] and is_selected[index]:
                        is_selected[index] = False
        res = results[f]
        keep = []
        for index in range(merged_count[f]):
            if not is_selected[index]:
                continue
            k = base + index
            if not bbox_ok[k]:
                bb = bbox[k]
                logger.warning("Invalid det bbox(%d,%d,%d,%d), skip it ..." % (bb[1], bb[3], bb[0], bb[2]))
                continue
            keep.append(k)
        res.class_ids_final = [merged_cls[k] for k in keep]
        res.class_names_final = [class_names[c] for c in res.class_ids_final]
        res.scores_final = [merged_score[k] for k in keep]
        res.bboxes = [bbox[k] for k in keep]
        res.captions = ["{} {:.2f}".format(n, sc) for n, sc in zip(res.class_names_final, res.scores_final)]
        final_planes.extend(keep)
        final_owner.extend([f] * len(keep))

    mark("host: cliques + selection")
    # -- final masks and pixel lists (make_json_results: np.argwhere(mask == 1), analyze.py:1903-1909)
    if final_planes:
        fin = ops.gather(planes, final_planes)
        unpacked = ops.unpack(fin, H, W) if want_masks else None
        owner = np.asarray(final_owner)
        fin_area = area[np.asarray(final_planes)]
        same_origin = all(tuple(o) == tuple(origins[0]) for o in origins)
        vx_all = None
        if same_origin:                                          # one launch and one copy for the whole batch
            px_all, off_all = ops.pixels(fin, H, W, fin_area, origins[0][0], origins[0][1])
            if vertexes is not None:                             # contours of every final object, one threaded C++ call
                vx_all = contours_of_flat_pixels(px_all, off_all, as_lists=(vertexes == "lists"))
        for f in range(F):
            rows = np.nonzero(owner == f)[0]
            if not len(rows):
                continue
            if same_origin:
                px, offsets = px_all, off_all[int(rows[0]):int(rows[-1]) + 2]
            else:                                                # a frame's final planes are contiguous
                px, offsets = ops.pixels(fin[int(rows[0]):int(rows[-1]) + 1], H, W, fin_area[rows], origins[f][0], origins[f][1])
            if vx_all is not None:
                results[f].vertexes = vx_all[int(rows[0]):int(rows[-1]) + 1]
            for k, row in enumerate(rows):
                results[f].pixels.append(px[offsets[k]:offsets[k + 1]])
                if want_masks:
                    full = unpacked[row]
                    results[f].masks_final.append(full.astype(np.int64) if merged_int[final_planes[row]] else full.view(np.bool_))
                else:
                    results[f].masks_final.append(None)
    mark("gpu+host: pixel lists / final masks")
    return results
