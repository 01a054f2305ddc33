"""bench.py --mode train: one training step of BASELINE.json configs[4] — random-init ResNet-101 Mask R-CNN, 2 images per GPU
at IMAGE_MAX_DIM = 256 with the SDetectorConfig training values (scripts/run.py:93-239), synthetic ground truth from the
generator's own source list, SGD momentum 0.9 / lr 5e-4 / clipnorm 5 — on N GPUs of one node, one process per GPU,
gradients averaged by the bucketed NCCL all-reduce of mrcnn/training.py (the replacement of mrcnn/parallel_model.py).

Prints ONE JSON line: images/s over all ranks (`value`: inputs resident in HBM; `e2e`: pinned host batch in, loss read back
every step), ms per step, the phase split (forward / backward / exposed all-reduce wait / optimiser), the all-reduce time of
the same buckets with nothing to hide behind, and the overlap derived from the two."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
S = 256


def _config():
    from mrcnn.config import Config

    class TrainConfig(Config):
        NAME = "rg-dataset"
        GPU_COUNT = 1
        IMAGES_PER_GPU = 2                      # configs[4]: nimg_per_gpu = 2
        NUM_CLASSES = 4
        IMAGE_MIN_DIM = S
        IMAGE_MAX_DIM = S
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        MAX_GT_INSTANCES = 300
        RPN_TRAIN_ANCHORS_PER_IMAGE = 512
        TRAIN_ROIS_PER_IMAGE = 512
        LEARNING_RATE = 0.0005
        USE_MINI_MASK = False
        DETECTION_MIN_CONFIDENCE = 0
    return TrainConfig()


def _dataset(first, count):
    import synth
    import torch
    from mrcnn import utils

    class SynthDataset(utils.Dataset):
        def __init__(self):
            super().__init__()
            for i, name in enumerate(["sidelobe", "source", "galaxy"]):
                self.add_class("rg", i + 1, name)
            self.items = {}
            for k in range(count):
                self.add_image("rg", image_id=first + k, path="synth://%d" % (first + k))

        def _item(self, image_id):
            gid = self.image_info[image_id]["id"]
            if gid not in self.items:
                m, masks, cls = synth.training_sample(gid, S)
                rgb = utils.maps_to_rgb8_device(torch.from_numpy(m[None]).cuda())[0]     # zscale stretch + uint8 RGB (csrc/preprocess.cu)
                self.items[gid] = (rgb[0].cpu().numpy(), masks, cls)
            return self.items[gid]

        def load_image(self, image_id):
            return self._item(image_id)[0]

        def load_mask(self, image_id):
            it = self._item(image_id)
            return it[1], it[2]

    ds = SynthDataset()
    ds.prepare()
    return ds


def run_train(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import synth
    from mrcnn import _native, training
    sys.path.insert(0, ROOT)
    from bench import ClockSampler

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = _config()
    B = cfg.BATCH_SIZE
    lib = _native.lib()
    graph = training.TrainGraph(cfg, device=local_rank, layers="all", seed=1000 * rank)
    graph.params.set_weights(synth.make_random_weights(0, 4))
    trainer = training.Trainer(graph, bucket_bytes=int(args.bucket_mb * 2 ** 20))
    trainer.broadcast_parameters()

    # distinct batches per rank through the real host data path (data_generator -> build_rpn_targets ...)
    n_sets = 4
    np.random.seed(1234 + rank)
    gen = training.data_generator(_dataset(rank * n_sets * B, n_sets * B), cfg, shuffle=False, batch_size=B)
    host_sets = []
    for _ in range(n_sets):
        inputs, _unused = next(gen)
        pinned = []
        for a in inputs:
            t = torch.from_numpy(np.ascontiguousarray(a))
            pinned.append((t.float() if t.dtype == torch.float64 else t).pin_memory())
        host_sets.append(pinned)
    dev_sets = [[t.cuda() for t in hs] for hs in host_sets]
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            fn(i)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    last = {}

    def dev_step(i):
        last.update(trainer.train_step(dev_sets[i % n_sets]))

    def e2e_step(i):
        dev = [t.to(graph.device, non_blocking=True) for t in host_sets[i % n_sets]]
        ls = trainer.train_step(dev)
        last.update(ls)
        last["loss_host"] = float(ls["loss"])            # device -> host read of the step's result

    graphed = False
    if not args.no_graph:
        graphed = trainer.capture(dev_sets[0])
    for i in range(args.warmup):
        dev_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = lib.mrcnn_kernel_launch_count()
    ms_total = timed(dev_step, args.steps)
    launches = lib.mrcnn_kernel_launch_count() - launches0
    clocks = sampler.stop()
    # phase split (forward / backward / exposed all-reduce wait / optimiser): a few eager steps with events between the
    # phases — a replayed graph has no seams to put events in
    saved_graph, trainer._graph = getattr(trainer, "_graph", None), None
    trainer.timing = True
    for i in range(4):
        dev_step(i)
    torch.cuda.synchronize()
    trainer.timing = False
    trainer._events = trainer._events[1:]
    phases = trainer.phase_ms()
    trainer._graph = saved_graph
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    # all-reduce cost: (a) the same buckets back to back with nothing to hide behind, (b) the step with the all-reduce
    # switched off (same capture mode) — what the step pays for the collective is T(with) - T(without), and the overlap
    # is the part of (a) that does not show up in it
    allreduce = None
    if world > 1:
        for _ in range(3):
            trainer.reducer.allreduce_only()
        reps = 10
        ms_ar = timed(lambda i: trainer.reducer.allreduce_only(), reps) / reps
        trainer.reducer.enabled = False
        if graphed:
            trainer.capture(dev_sets[0])
        for i in range(args.warmup):
            dev_step(i)
        ms_noar = timed(dev_step, args.steps) / args.steps
        trainer.reducer.enabled = True
        exposed = max(0.0, ms_total / args.steps - ms_noar)
        allreduce = {"bytes": int(graph.params.n) * 4, "buckets": len(trainer.reducer.buckets), "bucket_mb": args.bucket_mb,
                     "ms_alone": ms_ar, "ms_per_step_without_allreduce": ms_noar, "exposed_ms_per_step": exposed,
                     "overlap_pct": 100.0 * max(0.0, min(1.0, 1.0 - exposed / ms_ar)) if ms_ar > 0 else None,
                     "algbw_gbs": graph.params.n * 4 / (ms_ar / 1e3) / 1e9,
                     "eager_exposed_wait_ms": phases["allreduce_exposed"]}
    value = world * B * args.steps / (ms_total / 1e3)
    line = {"metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "mode": "train",
            "config": {"workload": "train mode ngpu=%d nimg_per_gpu=2, random-init ResNet-101 Mask R-CNN at IMAGE_MAX_DIM=256, "
                                   "NCCL gradient all-reduce (BASELINE.json configs[4])" % world,
                       "global_batch": world * B, "trainable_parameters": int(graph.params.trainable_elements),
                       "optimizer": "SGD momentum 0.9 lr 5e-4 clipnorm 5 + L2 1e-4/size", "train_rois_per_image": 512,
                       "dense_math": "torch bf16 conv/matmul (library); targets, ROIAlign fwd/bwd, proposals, optimiser: this repo's kernels",
                       "parallelism": "dp%d, one process per GPU" % world, "cuda_graph": bool(graphed)},
            "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
            "phase_ms_per_step_eager": phases, "allreduce": allreduce, "gpu_launches": int(launches), "clocks": clocks,
            "losses_last_step": {k: float(v.detach()) for k, v in last.items() if k != "loss_host"}, "grad_norm": trainer.opt.grad_norm()}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # a captured graph holds NCCL work: tearing the process group down while it is alive can block for minutes;
        # the measurement is done, leave without the interpreter's teardown
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)
