"""GPU parity of the tile driver: the pixel-list adjacency kernel, edge merging against the
REAL reference's outputs, and a whole tiled run against the same pipeline composed from per-tile pieces + the oracle."""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import sfinder_ops as S  # noqa: E402
from test_oracle_sfinder import summarise_sources  # noqa: E402
from test_sfinder_host import attach_tile_sources, make_finder  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sfinder_golden.json")


def write_fits(path, data):
    """Minimal FITS primary HDU, BITPIX -32."""
    cards = ["SIMPLE  =                    T", "BITPIX  =                  -32", "NAXIS   =                    2",
             "NAXIS1  = %20d" % data.shape[1], "NAXIS2  = %20d" % data.shape[0], "END"]
    header = "".join(c.ljust(80) for c in cards)
    header = header.ljust((len(header) + 2879) // 2880 * 2880)
    payload = np.ascontiguousarray(data, dtype=">f4").tobytes()
    payload += b"\0" * ((2880 - len(payload) % 2880) % 2880)
    with open(path, "wb") as f:
        f.write(header.encode("ascii") + payload)


def test_pixel_list_adjacency_kernel_matches_oracle():
    from mrcnn.sfinder import SFinder
    rng = np.random.default_rng(4)
    srcs = []
    for k in range(24):
        n = int(rng.choice([1, 3, 40, 700, 2500, 5000]))
        cy, cx = rng.integers(0, 90, size=2)
        pts = np.unique(np.stack([rng.integers(cy, cy + 60, size=n), rng.integers(cx, cx + 60, size=n)], axis=1), axis=0)
        srcs.append({"pixels": pts.tolist() if k % 2 else pts.astype(np.int32)})
    srcs.append({"pixels": [[500, 500]]})
    srcs.append({"pixels": [[501, 501]]})          # diagonal neighbour: adjacent (8-connectivity)
    srcs.append({"pixels": [[500, 502]]})          # two columns away: not adjacent
    pairs = [(i, j) for i in range(len(srcs)) for j in range(i + 1, len(srcs))]
    sf = SFinder(None, object())
    got = sf._adjacent_on_device(srcs, pairs)
    want = [S.pixels_adjacent(srcs[i]["pixels"], srcs[j]["pixels"]) for i, j in pairs]
    assert got == want
    assert 0 < sum(want) < len(want)


def test_merge_edge_sources_matches_reference_goldens():
    golden = json.load(open(GOLDEN))
    for case in golden["cases"]:
        sf = make_finder(case["setup"])
        attach_tile_sources(sf, case, range(case["setup"]["nproc"]))
        for w in range(case["setup"]["nproc"]):
            sf.procId = w
            for j in range(len(sf.tasks_per_worker[w])):
                sf.find_sources_at_edge(j)
        sf.procId = 0
        sf.tile_sources = {"sources": [t.det_sources for w in sf.tasks_per_worker for t in w if t.det_sources]}
        assert sf.merge_edge_sources() == 0
        assert summarise_sources(sf.sources["sources"]) == case["sources"], case["setup"]


def test_tiled_run_equals_per_tile_pipeline_plus_oracle_merge(tmp_path):
    """SFinder.run_parallel on a 2 x 3 tile grid (batched, overlapped, device-resident masks) == Analyzer.predict per
    tile on host arrays + the oracle's edge flagging / merging."""
    import synth
    from mrcnn import model as modellib, utils
    from mrcnn.analyze import Analyzer
    from mrcnn.sfinder import SFinder
    from oracle import network as N
    from test_gpu_engine import _config

    B = 2
    cfg = _config(B)
    cfg.CLASS_NAMES = ["bkg", "spurious", "compact", "extended"]
    big = np.concatenate([np.concatenate(list(synth.radio_maps(3, 100, start=10 * r)), axis=1) for r in range(2)], axis=0)
    big = big[:190, :280]                                   # ragged last row / column of tiles: 100x100, 100x80, 90x100, 90x80
    path = str(tmp_path / "mosaic.fits")
    write_fits(path, big)
    cfg.IMG_PATH = path
    cfg.SPLIT_IMG_IN_TILES, cfg.TILE_XSIZE, cfg.TILE_YSIZE, cfg.TILE_XSTEP, cfg.TILE_YSTEP = True, 100, 100, 1.0, 1.0
    cfg.ZSCALE_CONTRASTS = [0.25, 0.25, 0.25]
    cfg.IOU_THR = 0.6
    m = modellib.MaskRCNN(mode="inference", config=cfg, model_dir=str(tmp_path))
    m.set_weights(N.make_random_weights(0, 4))
    probe = m.detect_maps(np.stack([big[:100, :100], big[:100, 100:200]]).astype(np.float32))
    cfg.SCORE_THR = float(np.median(np.concatenate([r["scores"] for r in probe])))

    sf = SFinder(m, cfg)
    sf.outfile_json = str(tmp_path / "catalog.json")
    assert sf.run_parallel() == 0
    assert [t.coords for t in sf.tasks_per_worker[0]] == [(0, 100, 0, 100), (100, 200, 0, 100), (200, 280, 0, 100),
                                                         (0, 100, 100, 190), (100, 200, 100, 190), (200, 280, 100, 190)]

    # the same thing piece by piece: read_fits tile -> detect (batch filled with the tile) -> Analyzer on host arrays
    tiles = []
    for t in sf.tasks_per_worker[0]:
        image, _ = utils.read_fits(path, t.ix_min, t.ix_max, t.iy_min, t.iy_max, zscale_contrasts=cfg.ZSCALE_CONTRASTS)
        r = m.detect([image] * B)[0]
        an = Analyzer(m, cfg)
        an.class_names, an.score_thr, an.obj_name_tag = cfg.CLASS_NAMES, cfg.SCORE_THR, t.sname_tag
        an.image, an.image_id, an.image_xmin, an.image_ymin = image, sf.image_id, t.ix_min, t.iy_min
        an.masks, an.boxes, an.class_ids, an.scores = r["masks"], r["rois"], r["class_ids"], r["scores"]
        an.extract_det_masks()
        an.make_json_results()
        if an.results["objs"]:
            neighbors = [sf.tasks_per_worker[w][k].coords for w, k in zip(t.neighborWorkerId, t.neighborTaskIndex)]
            objs = copy.deepcopy(an.results["objs"])
            S.find_sources_at_edge(objs, t.coords, neighbors)
            tiles.append({"objs": objs, "workerId": 0, "tileId": t.tid, "neighborTileIds": t.neighborTaskId})
    want = S.merge_edge_sources(tiles)
    got = sf.sources["sources"]
    assert len(got) == len(want) > 0
    for s in got:
        s["vertexes"] = []
    for s in want:
        s["vertexes"] = []
    assert summarise_sources(got) == summarise_sources(want)
    assert any(s["edge"] for s in want)
    written = json.load(open(sf.outfile_json))
    assert [s["name"] for s in written["sources"]] == [s["name"] for s in want]
