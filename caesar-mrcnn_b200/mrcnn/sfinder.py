"""SFinder: source finding on a whole FITS image, serially or tile by tile (reference: mrcnn/sfinder.py).

Mirrors the detect side of the reference's `SFinder` / `TileTask`: `run()` (:450-526), `run_parallel()` (:549-640),
`create_tile_tasks()` (:1216-1384), `find_sources_at_edge()` (:643-706), `merge_edge_sources()` (:711-935),
`gather_task_data_from_workers()` (:940-988), `save()` / `write_json_results()` (:1389-1433), with the same attribute
names, tile / neighbour bookkeeping and catalogue layout.

B200 design. The reference runs one MPI process per group of tiles, each of which reads its tile, runs the detector
at batch 1 and post-processes on the host; the master then compares the pixel lists of all edge sources pairwise in
Python (an O(npix_a * npix_b) double loop per pair). Here one process per GPU takes the tiles the reference would give
to that rank, cuts them from the image once, and pushes them through the detector in batches of BATCH_SIZE with the
post-processing of batch k overlapping the detection of batch k+1 (`Analyzer.predict_maps_stream`; masks never leave
the GPU). Ranks exchange their per-tile catalogues with `torch.distributed.all_gather_object` (the reference's MPI
send/recv to the master) — there is no collective on the data path. The master's pair test runs as one kernel launch
over all candidate pairs (`mrcnn_pixel_lists_adjacent`, csrc/analyze.cu); the graph of mergeable sources, the order
of the merged pixel lists and the reference's choice of class / score for a merged source (it takes the LAST member
of the group: sfinder.py:862-864 index with the loop variable) are reproduced exactly.

Not provided: compute_source_params (WCS / flux parameters, needs astropy), DS9 regions, the "vertexes" of merged
sources without scikit-image (same rule as mrcnn/analyze.py). A non-MPI tile run of the reference never fills
`tile_sources` (gather is skipped, sfinder.py:600) and so writes an empty catalogue; here the local tiles are always
aggregated, i.e. a single process behaves like an MPI run with one rank.
"""
import ctypes
import json
import os

import numpy as np

from . import _native, fitsio, logger, utils
from .analyze import Analyzer, Graph, NumpyEncoder, contours_of_pixel_lists


class TileTask(object):
    """One tile of the image (reference: sfinder.py:54-166); coordinates (ix_min, ix_max, iy_min, iy_max), maxima
    exclusive as produced by utils.generate_tiles."""

    def __init__(self, tile_coords, model, config):
        self.model = model
        self.config = config
        self.coords = tile_coords
        self.ix_min, self.ix_max, self.iy_min, self.iy_max = tile_coords
        self.wid = -1
        self.tid = 0
        self.sname_tag = ""
        self.neighborTaskId = []
        self.neighborTaskIndex = []
        self.neighborWorkerId = []
        self.det_sources = {}
        self.save_json = False
        self.save_regions = False

    def set_worker_id(self, wid):
        self.wid = wid

    def set_task_id(self, tid):
        self.tid = tid
        self.sname_tag = "t" + str(tid)

    def is_task_tile_adjacent(self, aTask):
        in_x = (self.ix_max == aTask.ix_min - 1 or self.ix_min == aTask.ix_max + 1 or
                (self.ix_min == aTask.ix_min and self.ix_max == aTask.ix_max))
        in_y = (self.iy_max == aTask.iy_min - 1 or self.iy_min == aTask.iy_max + 1 or
                (self.iy_min == aTask.iy_min and self.iy_max == aTask.iy_max))
        return in_x and in_y

    def is_task_tile_overlapping(self, aTask):
        return not (self.ix_max < aTask.ix_min or self.ix_min > aTask.ix_max or
                    self.iy_max < aTask.iy_min or self.iy_min > aTask.iy_max)

    def is_task_tile_neighbor(self, aTask):
        return self.is_task_tile_adjacent(aTask) or self.is_task_tile_overlapping(aTask)

    def add_neighbor_info(self, tid, tindex, wid):
        self.neighborTaskId.append(tid)
        self.neighborTaskIndex.append(tindex)
        self.neighborWorkerId.append(wid)


class SFinder(object):
    """Source finder over one FITS image (config.IMG_PATH), reference: sfinder.py:264."""

    def __init__(self, model, config):
        self.config = config
        self.model = model
        self.header = None
        self.image_id = ""
        self.nx = -1
        self.ny = -1
        self.read_subimg = False
        self.xmin = self.xmax = self.ymin = self.ymax = -1
        self.tileSizeX = self.tileSizeY = -1
        self.tileStepSizeX = self.tileStepSizeY = 1
        self.mpiEnabled = False
        self.nproc = 1
        self.procId = 0
        self.MASTER_ID = 0
        self.tasks_per_worker = []
        self.tile_sources = {"sources": []}
        self.sources = {"sources": []}
        self.save_tile_json = False
        self.save_tile_regions = False
        self.write_to_json = True
        self.write_to_ds9 = False
        self.outfile_json = ""
        self.pixels_as_lists = True
        # results of the serial run
        self.bboxes_det = self.scores_det = self.classid_det = self.masks_det = None

    # -- workers = torch.distributed ranks (one process per GPU) ------------------------------------------
    def init_mpi(self):
        """Reference: MPI.COMM_WORLD (sfinder.py:528-546). Here: the torch.distributed world, when initialised."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.mpiEnabled = True
                self.nproc = dist.get_world_size()
                self.procId = dist.get_rank()
                return
        except ImportError:
            pass
        self.mpiEnabled, self.nproc, self.procId = False, 1, 0

    def set_img_size_params(self):
        """Reference: sfinder.py:336-444 (image range, tile geometry; beam / WCS parameters are not needed here)."""
        self.header = utils.get_fits_header(self.config.IMG_PATH)
        if self.header is None:
            logger.error("[PROC %d] Header read from image %s is None!" % (self.procId, self.config.IMG_PATH))
            return -1
        xmin, xmax = getattr(self.config, "IMG_XMIN", -1), getattr(self.config, "IMG_XMAX", -1)
        ymin, ymax = getattr(self.config, "IMG_YMIN", -1), getattr(self.config, "IMG_YMAX", -1)
        if xmin >= 0 and xmax >= 0 and ymin >= 0 and ymax >= 0:
            self.read_subimg = True
            self.nx = (self.xmax - self.xmin + 1)      # sic: computed from the previous range, as in the reference
            self.ny = (self.ymax - self.ymin + 1)
            self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax
        else:
            self.read_subimg = False
            if 'NAXIS1' not in self.header or 'NAXIS2' not in self.header:
                logger.error("[PROC %d] NAXIS1/NAXIS2 keyword missing in header!" % self.procId)
                return -1
            self.nx = self.header['NAXIS1']
            self.ny = self.header['NAXIS2']
            self.xmin, self.xmax, self.ymin, self.ymax = 0, self.nx - 1, 0, self.ny - 1
        self.tileSizeX, self.tileSizeY = self.nx, self.ny
        self.tileStepSizeX = self.tileStepSizeY = 1
        if getattr(self.config, "SPLIT_IMG_IN_TILES", False):
            self.tileSizeX, self.tileSizeY = self.config.TILE_XSIZE, self.config.TILE_YSIZE
            self.tileStepSizeX, self.tileStepSizeY = self.config.TILE_XSTEP, self.config.TILE_YSTEP
        self.image_id = os.path.splitext(os.path.basename(os.path.abspath(self.config.IMG_PATH)))[0]
        return 0

    # -- serial run ------------------------------------------------------------------------------
    def _analyzer(self):
        analyzer = Analyzer(self.model, self.config)
        analyzer.draw = False
        analyzer.write_to_ds9 = False
        analyzer.iou_thr = self.config.IOU_THR
        analyzer.score_thr = self.config.SCORE_THR
        return analyzer

    def run(self):
        """Reference: sfinder.py:450-526 (whole image or sub-image through Analyzer.predict)."""
        if self.set_img_size_params() < 0:
            logger.error("Failed to set image size parameters!")
            return -1
        if self.read_subimg:
            res = utils.read_fits(self.config.IMG_PATH, self.xmin, self.xmax, self.ymin, self.ymax,
                                  zscale_contrasts=self.config.ZSCALE_CONTRASTS)
        else:
            res = utils.read_fits(self.config.IMG_PATH, zscale_contrasts=self.config.ZSCALE_CONTRASTS)
        if res is None:
            logger.error("Failed to read image %s!" % self.config.IMG_PATH)
            return -1
        image_data, header = res
        analyzer = self._analyzer()
        analyzer.write_to_json = True
        analyzer.outfile_json = getattr(self.config, "OUTFILE_JSON", "")
        if analyzer.predict(image_data, self.image_id, header=header) < 0:
            logger.error("Failed to run model prediction on image %s!" % self.config.IMG_PATH)
            return -1
        self.bboxes_det, self.scores_det = analyzer.bboxes, analyzer.scores_final
        self.classid_det, self.masks_det = analyzer.class_ids_final, analyzer.masks_final
        if not self.bboxes_det:
            logger.info("No object detected in image %s ..." % self.config.IMG_PATH)
            return 0
        logger.info("#%d objects found in image %s ..." % (len(self.bboxes_det), self.config.IMG_PATH))
        return 0

    # -- tile run ----------------------------------------------------------------------------------
    def create_tile_tasks(self):
        """Reference: sfinder.py:1216-1384 — tiles, round-robin assignment to workers, neighbour lists."""
        tileGrid = utils.generate_tiles(self.xmin, self.xmax, self.ymin, self.ymax, self.tileSizeX, self.tileSizeY,
                                        self.tileStepSizeX, self.tileStepSizeY)
        if tileGrid is None:
            return -1
        self.tasks_per_worker = [[] for _ in range(self.nproc)]
        workerCounter = 0
        for tid, coords in enumerate(tileGrid):
            task = TileTask(coords, self.model, self.config)
            task.set_worker_id(workerCounter)
            task.set_task_id(tid)
            task.save_regions, task.save_json = self.save_tile_regions, self.save_tile_json
            self.tasks_per_worker[workerCounter].append(task)
            workerCounter = 0 if workerCounter >= self.nproc - 1 else workerCounter + 1
        for i, mine in enumerate(self.tasks_per_worker):
            for j, task in enumerate(mine):
                for k in range(j + 1, len(mine)):              # neighbours inside the same worker first
                    if task.is_task_tile_neighbor(mine[k]):
                        task.add_neighbor_info(mine[k].tid, k, i)
                        mine[k].add_neighbor_info(task.tid, j, i)
                for s in range(i + 1, len(self.tasks_per_worker)):      # then across workers
                    for t, other in enumerate(self.tasks_per_worker[s]):
                        if task.is_task_tile_neighbor(other):
                            task.add_neighbor_info(other.tid, t, s)
                            other.add_neighbor_info(task.tid, j, i)
        limit = getattr(self.config, "MAX_NTASKS_PER_WORKER", 100)
        if any(len(w) > limit for w in self.tasks_per_worker):
            logger.warning("[PROC %d] Too many tasks per worker exceeded (thr=%d)!" % (self.procId, limit))
            return -1
        return 0

    def _find_sources_in_my_tiles(self):
        """TileTask.find_sources (sfinder.py:169-262) for every tile of this rank, batched through the detector."""
        mine = self.tasks_per_worker[self.procId]
        if not mine:
            return 0
        try:
            data, _ = fitsio.read_primary(self.config.IMG_PATH)
        except Exception:
            logger.error("[PROC %d] Cannot read image file %s" % (self.procId, self.config.IMG_PATH))
            return -1
        plane = data[0, 0] if data.ndim == 4 else data
        B = self.config.BATCH_SIZE
        # the image goes to the GPU once (native-endian float32); tiles are cut there, on the detector's stream, so
        # the per-batch host work is a handful of slice objects instead of a 16 MB byte-swapping copy
        torch = utils._torch()
        stream = self.model._stream
        with torch.cuda.stream(stream):
            d_plane = torch.from_numpy(np.ascontiguousarray(plane, dtype=np.float32)).to("cuda:%d" % int(self.model._device))
        by_shape = {}
        for j, task in enumerate(mine):
            shape = (min(task.iy_max, d_plane.shape[0]) - task.iy_min, min(task.ix_max, d_plane.shape[1]) - task.ix_min)
            by_shape.setdefault(shape, []).append(j)

        def batches():
            for shape, items in by_shape.items():
                for lo in range(0, len(items), B):
                    chunk = items[lo:lo + B]
                    padded = chunk + [chunk[-1]] * (B - len(chunk))          # the tail batch repeats its last tile
                    tasks = [mine[j] for j in padded]
                    with torch.cuda.stream(stream):
                        maps = torch.stack([d_plane[t.iy_min:t.iy_max, t.ix_min:t.ix_max] for t in tasks]).contiguous()
                    order.append(list(chunk))
                    yield (maps, [self.image_id] * B, [(t.iy_min, t.ix_min) for t in tasks], [t.sname_tag for t in tasks])

        order = []
        analyzer = self._analyzer()
        analyzer.pixels_as_lists = self.pixels_as_lists
        for k, catalogues in enumerate(analyzer.predict_maps_stream(batches(), tuple(self.config.ZSCALE_CONTRASTS))):
            for j, cat in zip(order[k], catalogues):
                task = mine[j]
                if not cat["objs"]:
                    logger.info("[PROC %d] No object detected in tile image for task %d ..." % (self.procId, task.tid))
                    continue
                cat.update({"workerId": task.wid, "tileId": task.tid, "neighborTileIds": task.neighborTaskId,
                            "xmin": task.ix_min, "xmax": task.ix_max, "ymin": task.iy_min, "ymax": task.iy_max})
                task.det_sources = cat
                if task.save_json:
                    with open('catalog_' + self.image_id + '_tid' + str(task.tid) + '.json', 'w') as fp:
                        json.dump(cat, fp, indent=2, sort_keys=True, cls=NumpyEncoder)
        return 0

    def run_parallel(self):
        """Reference: sfinder.py:549-640, one process per GPU instead of one MPI rank per tile group."""
        self.init_mpi()
        if self.set_img_size_params() < 0:
            return -1
        if self.create_tile_tasks() < 0:
            logger.warning("[PROC %d] Failure in create tile tasks, exit..." % self.procId)
            return -1
        status = self._find_sources_in_my_tiles()
        for j in range(len(self.tasks_per_worker[self.procId])):
            self.find_sources_at_edge(j)
        if status < 0:
            logger.warning("[PROC %d] One or more errors occurred in source finding tasks..." % self.procId)
        if self.gather_task_data_from_workers() < 0:
            return -1
        if self.procId == self.MASTER_ID:
            self.merge_edge_sources()
            self.save()
        return 0

    def find_sources_at_edge(self, tindex):
        """Reference: sfinder.py:643-706 — flags sources on the tile border or inside a neighbour tile's range."""
        tileData = self.tasks_per_worker[self.procId][tindex]
        if not tileData.det_sources or not tileData.det_sources["objs"]:
            return
        xmin, xmax, ymin, ymax = tileData.ix_min, tileData.ix_max, tileData.iy_min, tileData.iy_max
        neighbors = [self.tasks_per_worker[w][t] for w, t in zip(tileData.neighborWorkerId, tileData.neighborTaskIndex)]
        for source in tileData.det_sources["objs"]:
            xmin_s, xmax_s, ymin_s, ymax_s = source["x1"], source["x2"], source["y1"], source["y2"]
            if xmin_s == xmin or xmax_s == xmax or ymin_s == ymin or ymax_s == ymax:
                source["edge"] = True
                continue
            for n in neighbors:
                if xmax_s < n.ix_min or xmin_s > n.ix_max or ymax_s < n.iy_min or ymin_s > n.iy_max:
                    continue
                source["edge"] = True
                break

    def gather_task_data_from_workers(self):
        """Reference: sfinder.py:940-988 (MPI send/recv to the master). Every rank contributes the catalogues of its
        tiles; the master receives them in worker order."""
        self.tile_sources = {"sources": [t.det_sources for t in self.tasks_per_worker[self.procId] if t.det_sources]}
        if not self.mpiEnabled:
            return 0
        import torch.distributed as dist
        gathered = [None] * self.nproc
        dist.all_gather_object(gathered, self.tile_sources)
        if self.procId == self.MASTER_ID:
            for i in range(1, self.nproc):
                if self.tasks_per_worker[i]:
                    self.tile_sources["sources"].extend(gathered[i]["sources"])
        return 0

    def merge_edge_sources(self):
        """Reference: sfinder.py:711-935. The pairwise pixel test runs on the GPU for all candidate pairs at once."""
        if self.procId != self.MASTER_ID:
            return 0
        tiles = self.tile_sources["sources"]
        self.sources["sources"] = []
        to_merge = []                                     # (sindex, tindex) of every edge source, in tile order
        for tindex, tileData in enumerate(tiles):
            for sindex, source in enumerate(tileData["objs"]):
                if not source["edge"]:
                    source["merged"] = False
                    self.sources["sources"].append(source)
                else:
                    to_merge.append((sindex, tindex))
        N = len(to_merge)
        srcs = [tiles[t]["objs"][s] for s, t in to_merge]
        # candidate pairs: the other source's tile is a neighbour of this one's and the bounding boxes overlap
        pairs = []
        if N > 1:
            box = np.array([[s["x1"], s["x2"], s["y1"], s["y2"]] for s in srcs], dtype=np.int64)
            tile_id = [tiles[t]["tileId"] for _, t in to_merge]
            neighbor_sets = [set(tiles[t]["neighborTileIds"]) for _, t in to_merge]
            for i in range(N - 1):
                j = np.arange(i + 1, N)
                ok = ~((box[i, 1] < box[j, 0]) | (box[i, 0] > box[j, 1]) | (box[i, 3] < box[j, 2]) | (box[i, 2] > box[j, 3]))
                for jj in j[ok]:
                    if tile_id[jj] in neighbor_sets[i]:
                        pairs.append((i, int(jj)))
        g = Graph(N)
        if pairs:
            for (i, j), hit in zip(pairs, self._adjacent_on_device(srcs, pairs)):
                if hit:
                    g.addEdge(i, j)
        for i, members in enumerate(g.connectedComponents()):
            sname_merged = "S" + str(i + 1) + "_merged"
            if len(members) == 1:
                source = srcs[members[0]]
                source["name"] = sname_merged
                source["merged"] = False
                self.sources["sources"].append(source)
                continue
            stacked = np.concatenate([np.asarray(srcs[m]["pixels"], dtype=np.int64).reshape(-1, 2) for m in members])
            # union without duplicates, first occurrence kept: the order of the reference's list concatenation
            key = stacked[:, 0] * (int(stacked[:, 1].max()) + 1) + stacked[:, 1]
            _, first = np.unique(key, return_index=True)
            pixels_merged = stacked[np.sort(first)]
            last = srcs[members[-1]]       # sic: the reference indexes with its loop variable, not with index_largest
            ymin, xmin = pixels_merged.min(axis=0)
            ymax, xmax = pixels_merged.max(axis=0)
            # contours of the merged mask (sfinder.py:885-910): the reference pads by 10 pixels and shifts back, which
            # gives the same image coordinates as the unit padding of the analyzer path
            vertex_list = contours_of_pixel_lists([pixels_merged])[0]
            self.sources["sources"].append({
                "name": sname_merged, "x1": xmin, "x2": xmax, "y1": ymin, "y2": ymax, "edge": True, "merged": True,
                "score": last["score"], "class_name": last["class_name"], "class_id": last["class_id"],
                "pixels": pixels_merged.tolist() if self.pixels_as_lists else pixels_merged.astype(np.int32),
                "vertexes": vertex_list})
        for i, source in enumerate(self.sources["sources"]):
            source["name"] = "S" + str(i + 1)
        return 0

    def _adjacent_on_device(self, srcs, pairs):
        """8-connected adjacency of the pixel lists of every candidate pair: one kernel launch (no CPU fallback)."""
        torch = utils._torch()
        lib = _native.lib()
        used = sorted({k for p in pairs for k in p})
        slot = {k: n for n, k in enumerate(used)}
        lists = [np.asarray(srcs[k]["pixels"], dtype=np.int32).reshape(-1, 2) for k in used]
        offsets = np.zeros(len(lists) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(a) for a in lists])
        device = "cuda:%d" % int(getattr(self.model, "_device", 0) or 0) if self.model is not None else "cuda:0"
        d_px = torch.from_numpy(np.ascontiguousarray(np.concatenate(lists))).to(device)
        d_off = torch.from_numpy(offsets).to(device)
        d_pairs = torch.from_numpy(np.array([[slot[i], slot[j]] for i, j in pairs], dtype=np.int32)).to(device)
        d_out = torch.empty((len(pairs),), dtype=torch.int32, device=device)
        with torch.cuda.device(d_out.device):
            _native.check(lib.mrcnn_pixel_lists_adjacent(_native.ptr(d_px), _native.ptr(d_off), _native.ptr(d_pairs), len(pairs),
                                                         _native.ptr(d_out), ctypes.c_void_p(torch.cuda.current_stream(d_out.device).cuda_stream)),
                          "pixel_lists_adjacent")
        return d_out.cpu().numpy().astype(bool).tolist()

    # -- output ------------------------------------------------------------------------------------
    def save(self):
        """Reference: sfinder.py:1389-1417 (JSON only)."""
        if self.procId != self.MASTER_ID:
            return
        if self.write_to_json:
            self.write_json_results(self.outfile_json if self.outfile_json != "" else 'catalog_' + str(self.image_id) + '.json')

    def write_json_results(self, outfile):
        """Reference: sfinder.py:1419-1433."""
        if self.procId != self.MASTER_ID:
            return
        if not self.sources:
            logger.warning("[PROC %d] Source dictionary is empty, nothing to be written ..." % self.procId)
            return
        with open(outfile, 'w') as fp:
            json.dump(self.sources, fp, indent=2, sort_keys=True, cls=NumpyEncoder)
