#!/usr/bin/env python
"""Times single layers of the bf16 implicit-GEMM convolution through the C ABI (mrcnn_conv2d_bf16) on their real shapes:
variants by environment (MRCNN_B200_BLOCK_N / _EPI_TMA / _OCC2), with and without the residual operand.
Development probe: back-to-back launches of one layer, CUDA events, median of 5 x 20 launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "caesar-mrcnn_b200"))
import torch
from mrcnn import _native as nat

LAYERS = {
    "res4_2a": dict(n=64, h=16, w=16, cin=1024, cout=256, k=1, relu=1, res=False),
    "res4_2b": dict(n=64, h=16, w=16, cin=256, cout=256, k=3, relu=1, res=False),
    "res4_2c": dict(n=64, h=16, w=16, cin=256, cout=1024, k=1, relu=1, res=True),
    "res3_2c": dict(n=64, h=32, w=32, cin=128, cout=512, k=1, relu=1, res=True),
    "res2_2c": dict(n=64, h=64, w=64, cin=64, cout=256, k=1, relu=1, res=True),
}


def time_layer(L, with_res, env):
    for k in ("MRCNN_B200_BLOCK_N", "MRCNN_B200_EPI_TMA", "MRCNN_B200_OCC2"):
        os.environ.pop(k, None)
    os.environ.update(env)
    lib = nat.lib()
    g = torch.Generator().manual_seed(1)
    x = torch.randn((L["n"], L["h"], L["w"], L["cin"]), generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn((L["cout"], L["k"] * L["k"] * L["cin"]), generator=g) * 0.05).to(torch.bfloat16).cuda()
    sc = torch.ones(L["cout"], device="cuda")
    sh = torch.zeros(L["cout"], device="cuda")
    res = torch.randn((L["n"], L["h"], L["w"], L["cout"]), generator=g).to(torch.bfloat16).cuda() if with_res else None
    out = torch.empty((L["n"], L["h"], L["w"], L["cout"]), dtype=torch.bfloat16, device="cuda")
    d = nat.ConvDesc(n=L["n"], h=L["h"], w=L["w"], cin=L["cin"], kh=L["k"], kw=L["k"], stride=1, pad=1 if L["k"] == 3 else 0,
                     cout=L["cout"], relu=L["relu"], residual_upsample2=0, out_dtype=1, out_mode=0, out_ld=0)
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        nat.check(lib.mrcnn_conv2d_bf16(d, nat.ptr(x), nat.ptr(w), nat.ptr(sc), nat.ptr(sh), nat.ptr(res), nat.ptr(out), st), "conv")
    for _ in range(5):
        launch()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 20 * 1e3)
    return sorted(times)[2]


if __name__ == "__main__":
    names = sys.argv[1:] or list(LAYERS)
    for name in names:
        L = LAYERS[name]
        for bn in ("128", "256"):
            for epi in ("1", "0"):
                for occ in ("0", "2"):
                    if occ == "2" and bn == "256":
                        continue
                    env = {"MRCNN_B200_BLOCK_N": bn, "MRCNN_B200_EPI_TMA": epi, "MRCNN_B200_OCC2": occ}
                    row = ["%-8s BN=%s epi_tma=%s occ2=%s" % (name, bn, epi, occ)]
                    for with_res in ([True, False] if L["res"] else [False]):
                        row.append("%s %.1f us" % ("res" if with_res else "nores", time_layer(L, with_res, env)))
                    print("  ".join(row), flush=True)
