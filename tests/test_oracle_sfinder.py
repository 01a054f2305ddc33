"""CPU: the tile-driver oracle (oracle/sfinder_ops.py) against outputs of the REAL reference
(tests/golden/make_golden_sfinder.py ran /root/reference/mrcnn/sfinder.py + utils.generate_tiles in the build
container)."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import sfinder_ops as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sfinder_golden.json")


@pytest.fixture(scope="module")
def golden_sf():
    with open(GOLDEN) as f:
        return json.load(f)


def summarise_sources(sources):
    out = []
    for s in sources:
        px = np.asarray(s["pixels"], dtype=np.int32).reshape(-1, 2)
        out.append({"name": s["name"], "x1": int(s["x1"]), "x2": int(s["x2"]), "y1": int(s["y1"]), "y2": int(s["y2"]),
                    "edge": bool(s["edge"]), "merged": bool(s["merged"]), "class_id": int(s["class_id"]), "class_name": s["class_name"],
                    "score_hex": float(s["score"]).hex(), "npix": int(len(px)),
                    "pixels_sha1": hashlib.sha1(np.ascontiguousarray(px).tobytes()).hexdigest()})
    return out


def tile_sources_in_gather_order(case, flag_edges):
    """Per-tile dicts in the MPI gather order (worker 0's tiles, worker 1's, ...), edge flags set by flag_edges."""
    tiles = []
    for worker in case["tasks"]:
        for t in worker:
            objs = copy.deepcopy(case["tile_objs"][str(t["tid"])])
            if not objs:
                continue
            neighbors = [case["tasks"][w][k]["coords"] for w, k in zip(t["neighborWorkerId"], t["neighborTaskIndex"])]
            flag_edges(objs, tuple(t["coords"]), [tuple(n) for n in neighbors])
            tiles.append({"objs": objs, "workerId": t["wid"], "tileId": t["tid"], "neighborTileIds": t["neighborTaskId"]})
    return tiles


def test_generate_tiles_matches_reference(golden_sf):
    assert len(golden_sf["tiles"]) >= 9
    for rec in golden_sf["tiles"]:
        grid = S.generate_tiles(*rec["args"])
        assert (None if grid is None else [list(t) for t in grid]) == rec["grid"], rec["args"]


def test_tile_tasks_edges_and_merging_match_reference(golden_sf):
    assert sum(sum(s["merged"] for s in c["sources"]) for c in golden_sf["cases"]) > 20
    for case in golden_sf["cases"]:
        su = case["setup"]
        grid = S.generate_tiles(0, su["nx"] - 1, 0, su["ny"] - 1, su["tile"][0], su["tile"][1], su["step"][0], su["step"][1])
        tasks = S.create_tile_tasks(grid, su["nproc"])
        assert [[{k: (list(t[k]) if k == "coords" else t[k]) for k in ("tid", "wid", "coords", "neighborTaskId", "neighborTaskIndex",
                                                                      "neighborWorkerId")} for t in w] for w in tasks] == case["tasks"]
        tiles = tile_sources_in_gather_order(case, S.find_sources_at_edge)
        assert {str(t["tileId"]): [bool(o["edge"]) for o in t["objs"]] for t in tiles} == case["edge_flags"]
        assert summarise_sources(S.merge_edge_sources(tiles)) == case["sources"]


def test_pixel_adjacency_is_8_connected():
    assert S.pixels_adjacent([[5, 5]], [[6, 6]]) and S.pixels_adjacent([[5, 5]], [[5, 5]])
    assert not S.pixels_adjacent([[5, 5]], [[7, 5]]) and not S.pixels_adjacent([[5, 5]], [[5, 7]])
    assert not S.pixels_adjacent([], [[1, 1]])
