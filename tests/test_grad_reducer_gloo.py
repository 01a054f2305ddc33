"""World-size-2 CPU (gloo) test of the training path's multi-GPU host logic: the bucketed gradient all-reduce of
mrcnn/training.py (GradReducer: buckets cut from the end of the flat buffer, launched from post-accumulate hooks while the
backward pass is still running, mean taken by the optimiser) must give every rank the gradient of the whole batch, and the
global-norm clip after the reduce must agree between ranks without another collective.
The graph is a small CPU test double with the same storage scheme (parameters = views of one flat buffer); the optimiser
kernel itself is CUDA-only and is covered by tests/test_gpu_training.py."""
import os
import sys
import tempfile

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Params(object):
    def __init__(self, sizes):
        self.order, self.entries, off = [], {}, 0
        for i, n in enumerate(sizes):
            self.order.append(("layer%d" % i, "kernel", off, n))
            self.entries[("layer%d" % i, "kernel")] = (off, (n,))
            off += (n + 7) // 8 * 8
        self.n = off
        self.w = torch.zeros(off)
        self.g = torch.zeros(off)


class _Graph(object):
    """y = sum_i tanh(x @ w_i[:d]) chained, parameters are views into params.w with .grad views into params.g"""
    def __init__(self, sizes, seed):
        self.params = _Params(sizes)
        gen = torch.Generator().manual_seed(seed)
        self.params.w.copy_(torch.randn(self.params.n, generator=gen) * 0.3)
        self.masters = {}
        for name, role, off, n in self.params.order:
            t = self.params.w[off:off + n].detach().requires_grad_(True)
            t.grad = self.params.g[off:off + n]
            self.masters[(name, role)] = t

    def loss(self, x):
        h = x
        for (name, role, off, n) in self.params.order:
            wv = self.masters[(name, role)]
            h = torch.tanh(h * wv[:h.shape[1]].view(1, -1) + wv.mean())
        return (h ** 2).mean()


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
    from mrcnn import training
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    sizes = [40, 300, 17, 1000, 64, 24]
    g = _Graph(sizes, seed=7)                                     # same weights on every rank
    red = training.GradReducer(g, bucket_bytes=4 * 100)           # ~100 elements per bucket -> several buckets
    gen = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 16, generator=gen)
    x = x_all[rank * 4:(rank + 1) * 4]
    for step in range(2):                                         # twice: the hook counters must re-arm
        g.params.g.zero_()
        g.loss(x).backward()
        red.finish()
    avg = g.params.g / world                                       # what mrcnn_sgd_step does with grad_scale = 1/world
    norm = float(avg.norm())
    np.save(os.path.join(out_dir, "grad_%d.npy" % rank), avg.numpy())
    np.save(os.path.join(out_dir, "meta_%d.npy" % rank), np.array([norm, len(red.buckets)]))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_equals_full_batch_gradient():
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, "init")
        mp.spawn(_worker, args=(2, init, d), nprocs=2, join=True)
        g0, g1 = np.load(os.path.join(d, "grad_0.npy")), np.load(os.path.join(d, "grad_1.npy"))
        m0, m1 = np.load(os.path.join(d, "meta_0.npy")), np.load(os.path.join(d, "meta_1.npy"))
    assert np.array_equal(g0, g1), "ranks disagree after the all-reduce"
    assert m0[0] == m1[0] and m0[1] >= 3                          # same clip norm everywhere, several buckets
    # single process, whole batch: mean over 8 samples == mean of the two 4-sample means
    ref = _Graph([40, 300, 17, 1000, 64, 24], seed=7)
    gen = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 16, generator=gen)
    ref.loss(x_all).backward()
    assert np.allclose(g0, ref.params.g.numpy(), rtol=1e-5, atol=1e-7)


def test_buckets_cover_the_buffer_once_from_the_end():
    sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
    from mrcnn import training
    g = _Graph([40, 300, 17, 1000, 64, 24], seed=1)
    red = training.GradReducer(g, bucket_bytes=4 * 100)
    assert red.world == 1
    spans = sorted(red.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == g.params.n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert red.buckets[0][1] == g.params.n                          # the first bucket is the END of the buffer (last layers)
    starts = {off for _, _, off, _ in g.params.order}
    assert all(s in starts for s, _ in red.buckets)                 # cut on tensor boundaries
