#!/usr/bin/env python
"""bench.py — detect images/s of the B200 Mask R-CNN detect path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over one batch of 64 synthetic FITS-like radio maps
 at IMAGE_MAX_DIM = 256: NaN fill + zscale + uint8 RGB -> resize/pad/mold ->
ResNet-101-FPN + RPN -> ProposalLayer -> ROIAlign -> class head -> DetectionLayer -> ROIAlign ->
mask head -> unmold to full-frame masks.  `value` times that with the float32 maps already resident
in HBM and results left in HBM; `e2e` times the same call with pinned HOST maps in and HOST results
(rois, class ids, scores, [H,W,N] masks) out.  Multi-GPU: one process per GPU, each rank runs its
own batches (weak scaling, no collective on the data path).

--impl reference times the CPU restatement of the reference path (oracle/: torch-CPU fp32 convs +
numpy/C++ layers; TensorFlow 1.13 cannot run in this image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))

import numpy as np  # noqa: E402

S = 256
BATCH = 64
METRIC = "detect_images_per_sec"
UNIT = "images/s"
WORKLOAD = "batched detect, 64 synthetic radio maps at IMAGE_MAX_DIM=256 (BASELINE.json configs[1])"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _oracle_cfg():
    return dict(PRE_NMS_LIMIT=6000, POST_NMS_ROIS_INFERENCE=1000, RPN_NMS_THRESHOLD=0.7,
                RPN_BBOX_STD_DEV=(0.1, 0.1, 0.2, 0.2), BBOX_STD_DEV=(0.1, 0.1, 0.2, 0.2), DETECTION_MIN_CONFIDENCE=0,
                DETECTION_NMS_THRESHOLD=0.3, DETECTION_MAX_INSTANCES=100, POOL_SIZE=7, MASK_POOL_SIZE=14)


def cpu_detect_images(maps, weights, threads, keep=None):
    """The reference detect path restated on the CPU (oracle): read_fits stretch -> mold -> graph -> unmold.
    keep: optional list that receives (detections [D,6], detect()-style dict) per image (parity check)."""
    from oracle import host_ops as H, network as N
    net = N.OracleNet(weights, 4, emulate_bf16=False, threads=threads)
    anchors = H.get_anchors((S, S, 3), (4, 8, 16, 32, 64))
    t0 = time.perf_counter()
    for m in maps:
        img = H.fits_to_rgb(m)
        molded, metas, windows = H.mold_inputs([img], min_dim=S, max_dim=S, min_scale=0, mode="square",
                                               mean_pixel=np.array([0, 0, 0]), num_classes=4)
        out = net.predict(molded, metas, anchors, _oracle_cfg())
        b, c, sc, mk = H.unmold_detections(out["detections"][0], out["mrcnn_mask"][0], img.shape, (S, S, 3), windows[0])
        if keep is not None:
            keep.append((out["detections"][0], {"rois": b, "class_ids": c, "scores": sc, "masks": mk}))
    return time.perf_counter() - t0


MAPS_TOTAL = 4096          # BASELINE.json configs[3]: the multi-GPU job is 4096 maps in contiguous shards


def job_maps(first, count):
    """Maps [first, first+count) of the 4096-map job: 256 generated radio maps x 16 rigid variants (8 dihedral x
    a half-frame roll) — all distinct frames with the same source statistics, without 68 s of map synthesis."""
    import synth
    base = {}
    out = np.empty((count, S, S), dtype=np.float32)
    for k in range(count):
        g = first + k
        b, v = g % 256, g // 256
        if b not in base:
            base[b] = synth.radio_map(b, S)
        m = base[b]
        if v & 1:
            m = m[::-1, :]
        if v & 2:
            m = m[:, ::-1]
        if v & 4:
            m = m.T
        if v & 8:
            m = np.roll(m, S // 2, axis=1)
        out[k] = m
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self._halt.is_set():
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._halt.set()
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.samples)}
        try:
            sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(self.samples[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, n in enumerate(names):
                if any(len(s) > 3 + k and s[3 + k].lower().startswith("active") for s in self.samples):
                    out["reasons"].append(n)
        except Exception:
            pass
        return out


def _bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) so that the
    pinned result buffers (420 MB per step and rank) are allocated on, and copied into, the GPU's own NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


def run_reference(args, rank, world):
    """Reference arm: CPU restatement on the host cores; rank 0 only."""
    if rank != 0:
        return
    import torch
    import synth
    threads = len(os.sched_getaffinity(0)) or 1
    torch.set_num_threads(threads)
    weights = synth.make_random_weights(0, 4)
    per_step = 1                                   # bounded sample: 1 image per step
    maps = synth.radio_maps(per_step * (args.steps + args.warmup), S)
    for w in range(args.warmup):
        cpu_detect_images(maps[w * per_step:(w + 1) * per_step], weights, threads)
    t = cpu_detect_images(maps[args.warmup * per_step:], weights, threads)
    n_img = per_step * args.steps
    value = n_img / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": per_step, "image_size": S},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d image(s) per step x %d steps, oracle CPU restatement (TF 1.13 cannot run here)" % (per_step, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    global S, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--image-size", type=int, default=S, help="IMAGE_MAX_DIM (256 = the run.py configuration the metric is "
                    "quoted on; 1024 = base Config stress size, use --batch 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="detect", choices=["detect", "train"], help="train: BASELINE.json configs[4] (train_bench.py)")
    ap.add_argument("--bucket-mb", type=float, default=25.0, help="--mode train: gradient all-reduce bucket size")
    ap.add_argument("--no-graph", action="store_true", help="--mode train: do not capture the step in a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.image_size != S or args.batch != BATCH:
        S = args.image_size
        WORKLOAD = "batched detect, %d synthetic radio maps at IMAGE_MAX_DIM=%d (secondary size; BASELINE.json configs[1] is 64 maps at 256)" % (args.batch, S)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "train":
        import train_bench
        train_bench.run_train(args, rank, world, local_rank)
        return

    # host threads of the result expansion: this rank's share of the cores (the ranks of one box share them)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if "MRCNN_B200_HOST_THREADS" not in os.environ:
        os.environ["MRCNN_B200_HOST_THREADS"] = str(max(1, min(16, (len(os.sched_getaffinity(0)) or 1) // max(1, local_world))))
    import torch
    import torch.distributed as dist
    import synth
    from mrcnn import _native, model as modellib
    from mrcnn.config import Config

    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = _bind_to_gpu_numa_node(local_rank)      # before any pinned allocation: first touch lands on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch

    class BenchConfig(Config):
        NAME = "rg-dataset"
        GPU_COUNT = 1
        IMAGES_PER_GPU = B
        NUM_CLASSES = 4
        IMAGE_MIN_DIM = S
        IMAGE_MAX_DIM = S
        RPN_ANCHOR_SCALES = (4, 8, 16, 32, 64)
        MEAN_PIXEL = np.array([0, 0, 0])
        DETECTION_MIN_CONFIDENCE = 0
        RPN_NMS_THRESHOLD = 0.7

    weights = synth.make_random_weights(0, 4)
    model = modellib.MaskRCNN(mode="inference", config=BenchConfig(), model_dir="/tmp/mrcnn_bench", device=local_rank)
    model.set_weights(weights)
    model.set_profiling(2)          # CUDA events where the kernel family changes (~35 per step), read after the timed loop
    lib = _native.lib()

    # distinct input batches per step so no step re-reads the previous step's inputs from L2;
    # the per-step working set (~12 GB of activations) is ~100x the 126 MB L2 anyway
    host_sets, dev_sets = [], []
    if world == 1:
        # configs[1]: 64 maps; 4 rotated / mirrored copies of the batch are cycled
        n_sets = 4
        base = synth.radio_maps(B, S, start=0)
        for k in range(n_sets):
            arr = np.roll(base, k * 7, axis=0).copy()
            if k % 2:
                arr = arr[:, ::-1, :].copy()
            host_sets.append(torch.from_numpy(arr).pin_memory())
        shard = (0, B)
    else:
        # configs[3]: rank r owns the contiguous shard [start, stop) of the 4096-map job and walks its batches
        # (cyclically when --steps exceeds the shard's batch count); no collective on the data path
        from mrcnn import sharding
        shard = sharding.shard_range(MAPS_TOTAL, rank, world)
        plan = [bt for bt in sharding.batches(shard[0], shard[1], B) if bt[1] == B]
        plan = plan[:max(4, min(len(plan), args.steps))]
        base = None
        for first, count in plan:
            arr = job_maps(first, count)
            if base is None:
                base = arr
            host_sets.append(torch.from_numpy(arr).pin_memory())
        n_sets = len(host_sets)
    for t in host_sets:
        dev_sets.append(t.cuda())
    stream = model._stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for i in range(steps):
                fn(i)
            ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- device-resident throughput ("value") ------------------------------------------------
    dev_step = lambda i: model.detect_maps(dev_sets[i % n_sets], device_only=True)  # noqa: E731
    for i in range(args.warmup):
        dev_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = lib.mrcnn_kernel_launch_count()
    ms_total = timed(dev_step, args.steps)
    launches = lib.mrcnn_kernel_launch_count() - launches0
    clocks = sampler.stop()
    # per-family device time: average of the last <= 8 timed steps, from the events recorded inside the timed region
    kt_acc = {k: [ms, n] for k, (ms, n) in model.kernel_times().items()}
    st_acc = model.stage_times()
    prof = {"n": 1}
    model.set_profiling(0)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end: pinned host maps in, host results out, every step -------------------------------
    # software-pipelined through the public API: batch i+1 is queued (H2D + compute) while the
    # device->host copy of batch i drains on the copy stream; every step's results are on the host
    # (and turned into detect()-style dicts) before the timed region ends
    e2e_results = [None]
    pending = [None]

    dbg = [] if os.environ.get("BENCH_DEBUG") else None

    def e2e_step(i):
        if dbg is not None:
            dbg.append(time.perf_counter())
        h = model.detect_maps_async(host_sets[i % n_sets])
        if pending[0] is not None:
            e2e_results[0] = pending[0].result()
        pending[0] = h
        if i == args.steps - 1:
            e2e_results[0] = pending[0].result()
            pending[0] = None
    # warm-up with the same retention pattern (3 result sets alive: retained, draining, in flight); the model's
    # pinned result pool is filled first (a fresh 420 MB cudaHostAlloc costs ~0.4 s)
    model.reserve_result_buffers(4, S, S)
    for i in range(max(args.warmup, 4)):
        h = model.detect_maps_async(host_sets[i % n_sets])
        if pending[0] is not None:
            e2e_results[0] = pending[0].result()
        pending[0] = h
    e2e_results[0] = pending[0].result()
    pending[0] = None
    ms_e2e = timed(e2e_step, args.steps)
    if dbg:
        sys.stderr.write("e2e host ms between steps: " + " ".join("%.1f" % (1e3 * (b - a)) for a, b in zip(dbg, dbg[1:])) + "\n")
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    dense_share = int(round((model._dense_share_state or {}).get("k", 0.0)))

    # the same pipeline with the masks left as the packed bits that crossed PCIe (result(expand=False), an extension of the
    # API): separates the device + PCIe path from the host-memory cost of materialising 6.5 MB of [H,W,N] bools per image
    def e2e_packed_step(i):
        h = model.detect_maps_async(host_sets[i % n_sets])
        if pending[0] is not None:
            e2e_results[1] = pending[0].result(expand=False)
        pending[0] = h
        if i == args.steps - 1:
            e2e_results[1] = pending[0].result(expand=False)
            pending[0] = None
    e2e_results.append(None)
    share_state = model._dense_share_state
    model._dense_share_state = {"fixed": True, "k": 0.0}      # packed results: nothing is delivered dense
    for i in range(3):
        e2e_packed_step(i)
    e2e_results[1] = pending[0].result(expand=False)
    pending[0] = None
    ms_e2e_packed = timed(e2e_packed_step, args.steps)
    model._dense_share_state = share_state
    D = 100
    h2d = B * S * S * 4 + B * 16 * 4 + B * 16
    # boxes / class ids / scores / counts + the pixel-major mask bits (16 B per pixel for D = 100); the [H,W,N] bool
    # arrays of the reference contract are expanded from the bits on the host inside the timed region
    d2h = B * (D * 4 * 4 + D * 4 + D * 4 + 4) + B * S * S * 4 * lib.mrcnn_mask_bits_words(D)

    # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions) ---------------
    peaks = _peaks()
    flops = model.flops_per_predict()
    gemm_ms, gemm_launches = kt_acc.get("conv_gemm", [0.0, 0])
    nprof = max(1, prof["n"])
    gemm_ms_step = gemm_ms / nprof
    achieved = flops / (gemm_ms_step / 1e3) / 1e12 if gemm_ms_step > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("conv_gemm_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel<BLOCK_N> (all %d launches of a step)" % (gemm_launches // nprof),
                "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic, "peak_source": peaks["source"] + ", sustained bf16",
                "flops_per_step": flops, "kernel_ms_per_step": gemm_ms_step}
    # HBM-bound stages the north star asks about (algorithmic bytes from SURVEY.md §8d, per image, bf16)
    alg = {"roialign": (27.89e6 + 12.82e6) * B, "proposal": 0.6707e6 * B, "detection": 0.096e6 * B}
    stages = {}
    for k, (ms, n) in kt_acc.items():
        per_step = ms / nprof
        ent = {"ms_per_step": per_step, "us_per_image": 1e3 * per_step / B, "launches_per_step": n // nprof}
        if k in alg and per_step > 0:
            gbs = alg[k] / (per_step / 1e3) / 1e9
            ent.update({"algorithmic_bytes_per_step": alg[k], "achieved_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]})
        stages[k] = ent

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "image_size": S, "num_classes": 4, "weights": "seeded random (LFS pointer unresolved)",
                       "l2": "%d distinct input batches cycled; per-step working set (activations) is GBs >> 126 MB L2" % n_sets,
                       "parallelism": "batch-sharded, no collective", "cpu_affinity": numa,
                       "maps_total": B if world == 1 else MAPS_TOTAL, "shard_of_rank0": list(shard),
                       "host_threads": int(lib.mrcnn_host_threads())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + dense_share * S * S * D,
                    "ms_per_step": ms_e2e / args.steps,
                    "masks": "shipped as pixel-major bits, expanded to [H,W,N] bool on the host (%d threads) inside the timed region"
                             % int(lib.mrcnn_host_threads()) + ("" if not dense_share else "; the masks of %d of the %d images per step "
                             "are expanded on the device instead and written dense by the DMA engine (share balanced against the host's "
                             "measured expansion rate, rank 0's value at the end of the run)" % (dense_share, B)),
                    "dense_share_images": dense_share},
            "e2e_packed_masks": {"value": world * B * args.steps / (ms_e2e_packed / 1e3), "unit": UNIT, "ms_per_step": ms_e2e_packed / args.steps,
                                 "note": "same calls, masks delivered to the host as packed bits (result(expand=False)); NOT the reference "
                                         "contract — shows what the [H,W,N] bool materialisation costs the host"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernel_families": stages,
            "stage_ms_per_step": {k: v / nprof for k, v in st_acc.items()},
            "detections_in_last_batch": int(sum(len(r["class_ids"]) for r in e2e_results[0]))}

    # ---- the same step when few of the 100 detection slots are used (N = 1 only; NOT the headline configuration) -------
    # Random weights fill every slot (DETECTION_MIN_CONFIDENCE = 0), the worst case the headline is quoted on.  A trained
    # model on a radio map fills a handful; the mask branch (45 % of the step) then runs on those only (the engine skips
    # the M tiles of zero-padded detections).  Second engine, same weights and maps, confidence cut at the 90th percentile.
    if world == 1 and os.environ.get("BENCH_SPARSE", "1") != "0":
        scores = np.concatenate([np.asarray(r["scores"]) for r in e2e_results[0]] + [np.zeros(0, np.float32)])
        if len(scores) >= 10:
            cut = float(np.quantile(scores, 0.9))

            class SparseConfig(BenchConfig):
                DETECTION_MIN_CONFIDENCE = cut

            sparse = modellib.MaskRCNN(mode="inference", config=SparseConfig(), model_dir="/tmp/mrcnn_bench", device=local_rank)
            sparse.set_weights(weights)
            sp_step = lambda i: sparse.detect_maps(dev_sets[i % n_sets], device_only=True)  # noqa: E731
            for i in range(args.warmup):
                sp_step(i)
            stream_main = stream
            stream = sparse._stream
            ms_sparse = timed(sp_step, args.steps)
            stream = stream_main
            n_det = sum(len(r["class_ids"]) for r in sparse.detect_maps(host_sets[0]))
            line["sparse_detections"] = {"min_confidence": cut, "detections_per_image": n_det / B, "ms_per_step": ms_sparse / args.steps,
                                         "value": B * args.steps / (ms_sparse / 1e3), "unit": UNIT,
                                         "note": "same maps and weights with DETECTION_MIN_CONFIDENCE at the 90th percentile of the "
                                                 "headline run's scores: tiles of zero-padded detections are skipped in the mask branch "
                                                 "(MRCNN_B200_SKIP_PADDED=0 computes them as the reference does); not the headline"}
            del sparse

    # ---- CPU baseline: bounded sample of the same workload on the host cores (rank 0, N = 1) --------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core again
        threads = len(all_cpus) or 1
        torch.set_num_threads(threads)
        sample = 4
        cpu_detect_images(base[:1], weights, threads)            # warm-up (page-in, thread pool)
        kept = []
        t = cpu_detect_images(base[:sample], weights, threads, keep=kept)
        line["cpu_baseline"] = {"value": sample / t, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d of the 64 maps, fp32 oracle restatement of the reference path (TF1 unavailable)" % sample}
        # the same 4 maps through the engine (bf16) against the fp32 oracle run just timed: final detections
        import parity_metrics as PM
        res = model.detect_maps(base)
        det = model.read_tensor("detections")
        m_det = [PM.match_detections_tensor(det[i], kept[i][0], S) for i in range(sample)]
        m_res = [PM.match_results(res[i], kept[i][1]) for i in range(sample)]
        sd, sr = PM.summarize(m_det), PM.summarize(m_res)
        line["parity_vs_fp32"] = {"images": sample, "matched_frac": sd["matched_frac"], "dbox_px_median": sd.get("dbox_px_median"),
                                  "dbox_px_p95": sd.get("dbox_px_p95"), "dscore_median": sd.get("dscore_median"),
                                  "dscore_p95": sd.get("dscore_p95"), "mask_iou_median": sr.get("mask_iou_median"),
                                  "mask_iou_ge_0.9_frac": sr.get("mask_iou_ge_0.9_frac"),
                                  "note": "bf16 engine vs fp32 CPU oracle, same maps and weights (tests/test_gpu_e2e_parity.py asserts the bounds)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
