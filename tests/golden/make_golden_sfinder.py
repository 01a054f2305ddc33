#!/usr/bin/env python
"""Generates tests/golden/sfinder_golden.json by running the REAL reference tile bookkeeping and edge merging.

Runs only in the build container (needs /root/reference). mrcnn/sfinder.py and mrcnn/utils.py are imported from the
reference with inert stubs for the absent packages (TensorFlow, Keras, astropy, scikit-image, OpenCV, regions,
numpyencoder, ...); skimage.measure.find_contours is a stub returning no contours, so "vertexes" is NOT pinned.
Executed reference code: utils.generate_tiles (utils.py:1254-1328), SFinder.create_tile_tasks (sfinder.py:1216-1384,
with TileTask's neighbour predicates :119-166), SFinder.find_sources_at_edge (:643-706), SFinder.merge_edge_sources
(:711-935). The per-tile source lists are synthetic: a seeded blob image is cut into the tiles the reference
generated and every 4-connected island inside a tile becomes one source (pixels in global coordinates, exclusive
x2 / y2 like Analyzer.make_json_results), which is what TileTask.find_sources hands to the merger.
Nothing here is imported by the product."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_from_reference as base  # noqa: E402

OUT = os.path.join(HERE, "sfinder_golden.json")
CLASS_NAMES = ["bkg", "spurious", "compact", "extended"]


def blob_image(rng, ny, nx, n_blobs):
    img = np.zeros((ny, nx), dtype=bool)
    yy, xx = np.mgrid[0:ny, 0:nx]
    for _ in range(n_blobs):
        cy, cx = int(rng.integers(0, ny)), int(rng.integers(0, nx))
        if rng.random() < 0.5:
            r = int(rng.integers(1, 7))
            img |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        else:
            img[cy:cy + int(rng.integers(1, 9)), cx:cx + int(rng.integers(1, 14))] = True
    return img


def tile_sources_from_image(img, tile, tid, rng):
    """Every 4-connected island of img inside tile (xmin, xmax, ymin, ymax; maxima exclusive) -> one source dict."""
    from scipy import ndimage
    xmin, xmax, ymin, ymax = tile
    sub = img[ymin:ymax, xmin:xmax]
    labels, n = ndimage.label(sub)
    objs = []
    for c in range(1, n + 1):
        px = np.argwhere(labels == c)
        class_id = int(rng.integers(1, len(CLASS_NAMES)))
        ny_t, nx_t = sub.shape
        y1, x1, y2, x2 = int(px[:, 0].min()), int(px[:, 1].min()), int(px[:, 0].max()) + 1, int(px[:, 1].max()) + 1
        at_edge = x1 <= 0 or x1 >= nx_t - 1 or x2 <= 0 or x2 >= nx_t - 1 or y1 <= 0 or y1 >= ny_t - 1 or y2 <= 0 or y2 >= ny_t - 1
        objs.append({"name": "S%d_t%d" % (c, tid), "x1": xmin + x1, "x2": xmin + x2, "y1": ymin + y1, "y2": ymin + y2,
                     "class_id": class_id, "class_name": CLASS_NAMES[class_id],
                     "score": float(np.float32(rng.uniform(0.7, 1.0))), "pixels": (px + [ymin, xmin]).tolist(), "vertexes": [],
                     "edge": bool(at_edge)})
    return objs


def summarise_sources(sources):
    out = []
    for s in sources:
        px = np.asarray(s["pixels"], dtype=np.int32).reshape(-1, 2)
        out.append({"name": s["name"], "x1": int(s["x1"]), "x2": int(s["x2"]), "y1": int(s["y1"]), "y2": int(s["y2"]),
                    "edge": bool(s["edge"]), "merged": bool(s["merged"]), "class_id": int(s["class_id"]), "class_name": s["class_name"],
                    "score_hex": float(s["score"]).hex(), "npix": int(len(px)),
                    "pixels_sha1": hashlib.sha1(np.ascontiguousarray(px).tobytes()).hexdigest()})
    return out


def main():
    base.install_stubs()
    for name in ("cv2", "imutils", "regions", "skimage.draw", "numpyencoder", "astropy.wcs.utils", "sklearn_stub_unused"):
        if name not in sys.modules:
            base._stub(name)
    sys.modules["skimage.measure"].find_contours = lambda image, level: []
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.path.insert(0, base.REF)
    import logging
    logging.disable(logging.CRITICAL)
    import contextlib
    import io
    from mrcnn import sfinder as ref_sfinder, utils as ref_utils

    golden = {"numpy": np.__version__, "class_names": CLASS_NAMES, "tiles": [], "cases": []}

    # ---- generate_tiles ------------------------------------------------------------------------
    for args in [(0, 255, 0, 255, 128, 128, 1.0, 1.0), (0, 299, 0, 199, 128, 96, 1.0, 1.0), (0, 299, 0, 199, 128, 96, 0.5, 0.75),
                 (10, 265, 20, 147, 64, 64, 1.0, 0.5), (0, 99, 0, 99, 100, 100, 1.0, 1.0), (0, 99, 0, 99, 101, 50, 1.0, 1.0),
                 (0, 99, 0, 99, 50, 50, 0.0, 1.0), (5, 5, 0, 9, 2, 2, 1.0, 1.0), (0, 511, 0, 383, 200, 150, 0.9, 0.33)]:
        note = ""
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                grid = ref_utils.generate_tiles(*args)
        except NameError as e:
            # the reference's utils.py never defines `logger`: every error branch of generate_tiles raises NameError
            # instead of logging and returning None; the intended result (None) is recorded
            grid, note = None, "reference raises NameError(%s) on its logging call; intended result None" % e
        golden["tiles"].append({"args": list(args), "grid": None if grid is None else [list(map(int, t)) for t in grid],
                                "note": note})

    # ---- create_tile_tasks + find_sources_at_edge + merge_edge_sources --------------------------------
    class Cfg:
        IMG_PATH = "/tmp/synthetic.fits"
        MPI = None
        MAX_NTASKS_PER_WORKER = 100
        NUM_CLASSES = len(CLASS_NAMES)

    rng = np.random.default_rng(20261019)
    setups = [dict(nx=256, ny=256, tile=(128, 128), step=(1.0, 1.0), nproc=1, blobs=40),
              dict(nx=300, ny=200, tile=(128, 96), step=(1.0, 1.0), nproc=3, blobs=60),
              dict(nx=300, ny=200, tile=(128, 96), step=(0.5, 0.75), nproc=2, blobs=45),
              dict(nx=192, ny=192, tile=(64, 64), step=(1.0, 1.0), nproc=4, blobs=70),
              dict(nx=160, ny=120, tile=(80, 60), step=(0.8, 0.8), nproc=8, blobs=30),
              dict(nx=128, ny=128, tile=(128, 128), step=(1.0, 1.0), nproc=2, blobs=12)]
    for setup in setups:
        nx, ny, nproc = setup["nx"], setup["ny"], setup["nproc"]
        img = blob_image(rng, ny, nx, setup["blobs"])
        sf = ref_sfinder.SFinder(None, Cfg())
        sf.xmin, sf.xmax, sf.ymin, sf.ymax = 0, nx - 1, 0, ny - 1
        sf.tileSizeX, sf.tileSizeY = setup["tile"]
        sf.tileStepSizeX, sf.tileStepSizeY = setup["step"]
        sf.nproc, sf.procId, sf.mpiEnabled = nproc, 0, False
        with contextlib.redirect_stdout(io.StringIO()):
            assert sf.create_tile_tasks() == 0
        tasks = [[dict(tid=t.tid, wid=t.wid, coords=[int(v) for v in t.coords], neighborTaskId=list(t.neighborTaskId),
                       neighborTaskIndex=list(t.neighborTaskIndex), neighborWorkerId=list(t.neighborWorkerId))
                  for t in worker] for worker in sf.tasks_per_worker]
        # per-tile sources (what TileTask.find_sources leaves in det_sources)
        inputs = {}
        for worker in sf.tasks_per_worker:
            for t in worker:
                objs = tile_sources_from_image(img, t.coords, t.tid, rng)
                inputs[str(t.tid)] = json.loads(json.dumps(objs))
                if objs:
                    t.det_sources = {"image_id": "synthetic", "objs": objs, "workerId": t.wid, "tileId": t.tid,
                                     "neighborTileIds": t.neighborTaskId, "xmin": t.ix_min, "xmax": t.ix_max,
                                     "ymin": t.iy_min, "ymax": t.iy_max}
        # every worker flags its own tiles, then the MPI gather order: worker 0's tiles, worker 1's, ...
        with contextlib.redirect_stdout(io.StringIO()):
            for w in range(nproc):
                sf.procId = w
                for j in range(len(sf.tasks_per_worker[w])):
                    sf.find_sources_at_edge(j)
            sf.procId = 0
            sf.tile_sources = {"sources": [t.det_sources for worker in sf.tasks_per_worker for t in worker if t.det_sources]}
            edge_flags = {str(ts["tileId"]): [bool(o["edge"]) for o in ts["objs"]] for ts in sf.tile_sources["sources"]}
            sf.merge_edge_sources()
        golden["cases"].append({"setup": {k: (list(v) if isinstance(v, tuple) else v) for k, v in setup.items()},
                                "image_rows_hex": [np.packbits(row).tobytes().hex() for row in img],
                                "tasks": tasks, "tile_objs": inputs, "edge_flags": edge_flags,
                                "sources": summarise_sources(sf.sources["sources"])})
    with open(OUT, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [len(c["sources"]) for c in golden["cases"]], "final sources;",
          [sum(s["merged"] for s in c["sources"]) for c in golden["cases"]], "merged")


if __name__ == "__main__":
    main()
