"""CPU: host logic of the tile driver (mrcnn/sfinder.py + utils.generate_tiles) against outputs of the REAL reference
(tests/golden/sfinder_golden.json) and, over gloo with world_size 2, the rank-to-master catalogue gather."""
import copy
import json
import os
import sys

import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "sfinder_golden.json")


@pytest.fixture(scope="module")
def golden_sf():
    with open(GOLDEN) as f:
        return json.load(f)


class _Cfg:
    IMG_PATH = "/tmp/synthetic.fits"
    MAX_NTASKS_PER_WORKER = 100


def make_finder(setup, proc_id=0):
    from mrcnn.sfinder import SFinder
    sf = SFinder(None, _Cfg())
    sf.xmin, sf.xmax, sf.ymin, sf.ymax = 0, setup["nx"] - 1, 0, setup["ny"] - 1
    sf.tileSizeX, sf.tileSizeY = setup["tile"]
    sf.tileStepSizeX, sf.tileStepSizeY = setup["step"]
    sf.nproc, sf.procId = setup["nproc"], proc_id
    assert sf.create_tile_tasks() == 0
    return sf


def attach_tile_sources(sf, case, workers):
    """What TileTask.find_sources leaves behind, for the tiles of the given workers."""
    for w in workers:
        for t in sf.tasks_per_worker[w]:
            objs = copy.deepcopy(case["tile_objs"][str(t.tid)])
            if objs:
                t.det_sources = {"image_id": "synthetic", "objs": objs, "workerId": t.wid, "tileId": t.tid,
                                 "neighborTileIds": t.neighborTaskId, "xmin": t.ix_min, "xmax": t.ix_max, "ymin": t.iy_min,
                                 "ymax": t.iy_max}


def test_generate_tiles_matches_reference(golden_sf):
    from mrcnn import utils
    for rec in golden_sf["tiles"]:
        grid = utils.generate_tiles(*rec["args"])
        assert (None if grid is None else [list(t) for t in grid]) == rec["grid"], rec["args"]
    assert utils.generate_tiles(0, 99, 0, 99, 50, 50, 0.001, 1.0) is None       # step rounds to 0: the reference never returns


def test_tile_tasks_and_edge_flags_match_reference(golden_sf):
    for case in golden_sf["cases"]:
        sf = make_finder(case["setup"])
        got = [[dict(tid=t.tid, wid=t.wid, coords=list(t.coords), neighborTaskId=t.neighborTaskId,
                     neighborTaskIndex=t.neighborTaskIndex, neighborWorkerId=t.neighborWorkerId) for t in w] for w in sf.tasks_per_worker]
        assert got == case["tasks"]
        attach_tile_sources(sf, case, range(case["setup"]["nproc"]))
        flags = {}
        for w in range(case["setup"]["nproc"]):
            sf.procId = w
            for j, t in enumerate(sf.tasks_per_worker[w]):
                sf.find_sources_at_edge(j)
                if t.det_sources:
                    flags[str(t.tid)] = [bool(o["edge"]) for o in t.det_sources["objs"]]
        assert flags == case["edge_flags"]


def _gather_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "caesar-mrcnn_b200"))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = json.load(open(GOLDEN))["cases"][2]          # 2 workers, overlapping tiles
    from test_sfinder_host import attach_tile_sources, make_finder
    sf = make_finder(case["setup"], rank)
    sf.init_mpi()
    assert (sf.mpiEnabled, sf.nproc, sf.procId) == (True, world, rank)
    attach_tile_sources(sf, case, [rank])                # every rank only knows the sources of its own tiles
    for j in range(len(sf.tasks_per_worker[rank])):
        sf.find_sources_at_edge(j)
    assert sf.gather_task_data_from_workers() == 0
    with open(os.path.join(out_dir, "r%d.json" % rank), "w") as f:
        json.dump([[t["tileId"], [bool(o["edge"]) for o in t["objs"]]] for t in sf.tile_sources["sources"]], f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_collects_tiles_in_worker_order(tmp_path, golden_sf):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    case = golden_sf["cases"][2]
    master = json.load(open(tmp_path / "r0.json"))
    want = [[t["tid"], case["edge_flags"][str(t["tid"])]] for w in case["tasks"] for t in w if case["tile_objs"][str(t["tid"])]]
    assert master == want                                 # worker 0's tiles, then worker 1's, edge flags included
    other = json.load(open(tmp_path / "r1.json"))
    assert [t[0] for t in other] == [t["tid"] for t in case["tasks"][1] if case["tile_objs"][str(t["tid"])]]
