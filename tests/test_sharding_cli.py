"""CPU: sharded-run host logic (gloo, world_size 2) and the CLI's argument handling."""
import os
import sys

import pytest
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200"))
    from mrcnn import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4096 + 3
    start, stop = sharding.shard_range(n, rank, world)
    done = sum(c for _, c in sharding.batches(start, stop, 64))
    counts = sharding.gather_counts(done)
    tmax = sharding.max_over_ranks(10.0 + rank)
    dist.barrier()
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write("%d %d %s %.1f" % (start, stop, ",".join(map(str, counts)), tmax))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything_once():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200"))
    from mrcnn import sharding
    for n in (0, 1, 63, 64, 4096, 4099):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(n, r, world)
                seen += list(range(a, b))
            assert seen == list(range(n))
    assert sharding.batches(10, 150, 64) == [(10, 64), (74, 64), (138, 12)]
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_two_rank_gloo_run(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = open(tmp_path / "r0.txt").read().split()
    r1 = open(tmp_path / "r1.txt").read().split()
    assert (int(r0[0]), int(r0[1]), int(r1[0]), int(r1[1])) == (0, 2050, 2050, 4099)
    assert r0[2] == r1[2] == "2050,2049"
    assert r0[3] == r1[3] == "11.0"          # max over ranks, identical everywhere


def test_cli_argument_validation(tmp_path, golden_dir):
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200", "scripts", "run.py")
    spec = importlib.util.spec_from_file_location("b200_run", path)
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    fits = os.path.join(golden_dir, "galaxy0002.fits")
    ok = run.parse_args(["detect", "--image", fits, "--random_weights", "0"])
    assert run.validate_args(ok) == 0
    cfg = run.make_config(ok)
    assert cfg.NUM_CLASSES == 4 and cfg.CLASS_NAMES == ["bkg", "sidelobe", "source", "galaxy"]
    assert cfg.IMAGE_META_SIZE == 16 and list(cfg.IMAGE_SHAPE) == [256, 256, 3] and cfg.BATCH_SIZE == 1
    assert cfg.RPN_ANCHOR_SCALES == (4, 8, 16, 32, 64) and cfg.DETECTION_MIN_CONFIDENCE == 0 and cfg.RPN_NMS_THRESHOLD == 0.7
    assert run.main(["train"]) == 1
    assert run.main(["detect", "--image", "/nonexistent.fits", "--weights", "w.h5"]) == 1
    assert run.main(["detect", "--image", fits]) == 1                              # no weights
    assert run.main(["detect", "--image", fits, "--random_weights", "0", "--grayimg"]) == 1
    assert run.main(["detect", "--image", fits, "--random_weights", "0", "--backbone", "resnet50"]) == 1
    assert run.main(["test", "--weights", "w.h5"]) == 1                            # no datalist
    # every flag of the reference's parser (scripts/run.py:1263-1384) is accepted, so command lines written for it still parse
    full = run.parse_args(["detect", "--image", fits, "--random_weights", "0", "--rpn_class_loss", "--no_mrcnn_mask_loss",
                           "--mrcnn_bbox_loss_weight", "2.5", "--mask_loss_function", "binary_crossentropy", "--bias", "0.4",
                           "--contrast", "1.2", "--classid_remap_dict", "{1:2}", "--datadir", "/data",
                           "--no_consider_sources_near_mixed_sidelobes", "--detect_outfile", "plot.png"])
    assert run.validate_args(full) == 0
    assert full.rpn_class_loss is True and full.mrcnn_mask_loss is False and full.mrcnn_class_loss is True
    assert full.mrcnn_bbox_loss_weight == 2.5 and full.bias == 0.4 and full.consider_sources_near_mixed_sidelobes is False
    assert run.main(["detect", "--image", fits, "--random_weights", "0", "--remap_classids"]) == 1        # empty dictionary
    assert run.main(["detect", "--image", fits, "--random_weights", "0", "--remap_classids", "--classid_remap_dict", "{1:2}"]) == 1
    ref_flags_path = "/root/reference/scripts/run.py"
    if os.path.exists(ref_flags_path):                                             # build container only: nothing is missing
        import re
        ref = set(re.findall(r"add_argument\('(--[A-Za-z_]+)'", open(ref_flags_path).read()))
        mine = set(a.option_strings[0] for a in run.build_parser()._actions if a.option_strings) if hasattr(run, "build_parser") else None
        if mine is not None:
            assert not (ref - mine), sorted(ref - mine)


def test_cli_tile_options_reach_the_tile_driver(golden_dir):
    """--split_img_in_tiles and the tile geometry flags land in the config and produce the reference's tile grid
    for the shipped 132 x 132 image (header-only FITS access: no GPU involved)."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "caesar-mrcnn_b200", "scripts", "run.py")
    spec = importlib.util.spec_from_file_location("b200_run_tiles", path)
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    fits = os.path.join(golden_dir, "galaxy0002.fits")
    args = run.parse_args(["detect", "--image", fits, "--random_weights", "0", "--split_img_in_tiles", "--tile_xsize", "66",
                           "--tile_ysize", "50", "--tile_xstep", "1.0", "--tile_ystep", "0.5", "--nimg_per_gpu", "4"])
    assert run.validate_args(args) == 0
    cfg = run.make_config(args)
    assert (cfg.SPLIT_IMG_IN_TILES, cfg.TILE_XSIZE, cfg.TILE_YSIZE, cfg.TILE_XSTEP, cfg.TILE_YSTEP) == (True, 66, 50, 1.0, 0.5)
    assert cfg.IMG_PATH == fits and cfg.BATCH_SIZE == 4
    sf = run.SFinder(None, cfg)
    assert sf.set_img_size_params() == 0
    assert (sf.nx, sf.ny, sf.xmin, sf.xmax, sf.ymin, sf.ymax, sf.image_id) == (132, 132, 0, 131, 0, 131, "galaxy0002")
    assert sf.create_tile_tasks() == 0
    coords = [t.coords for t in sf.tasks_per_worker[0]]
    assert coords[:2] == [(0, 66, 0, 50), (66, 132, 0, 50)] and coords[-1] == (66, 132, 125, 132) and len(coords) == 12
    assert sf.tasks_per_worker[0][0].neighborTaskId[:2] == [1, 2]
