"""GPU: the tcgen05/TMEM/TMA implicit-GEMM convolution (through the C ABI) against
(a) torch fp32 conv2d on the same bf16-rounded operands and (b) the CUDA-core reference kernel.
Floating-point stage: tolerance = bf16 output rounding (2^-8 relative) + fp32 accumulation-order
noise, written per assertion below."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import torch.nn.functional as F  # noqa: E402


def _native():
    from mrcnn import _native
    return _native


def run_conv(x, w2, scale, shift, residual, desc_kw, out_shape, out_dtype, simt=False):
    nat = _native()
    lib = nat.lib()
    d = nat.ConvDesc(**desc_kw)
    out = torch.full(out_shape, float("nan"), dtype=out_dtype, device="cuda")
    fn = lib.mrcnn_conv2d_bf16_simt if simt else lib.mrcnn_conv2d_bf16
    st = fn(d, nat.ptr(x), nat.ptr(w2), nat.ptr(scale), nat.ptr(shift), nat.ptr(residual), nat.ptr(out), None)
    nat.check(st, "conv2d")
    torch.cuda.synchronize()
    return out


def make_case(seed, n, h, w, cin, cout, k, stride=1, relu=False, residual=False, res_up2=False, out_f32=False,
              deconv=False, out_ld=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn((n, h, w, cin), generator=g) * 1.0).to(torch.bfloat16).cuda()
    pad = 1 if k == 3 else 0
    taps = 4 if deconv else 1
    wk = (torch.randn((taps * cout, k, k, cin), generator=g) / np.sqrt(k * k * cin)).to(torch.bfloat16).cuda()
    scale = (torch.rand((cout,), generator=g) + 0.5).cuda()
    shift = torch.randn((cout,), generator=g).cuda()
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    res = None
    if residual:
        rs = (n, oh // 2, ow // 2, cout) if res_up2 else (n, oh, ow, cout)
        res = torch.randn(rs, generator=g).to(torch.bfloat16).cuda()
    # fp32 reference on the same bf16 operands
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xf = x.float().permute(0, 3, 1, 2)
    wf = wk.float().permute(0, 3, 1, 2)
    acc = F.conv2d(xf, wf, None, stride=stride, padding=pad)          # [n, taps*cout, oh, ow]
    if deconv:
        acc = acc.reshape(n, 2, 2, cout, oh, ow)                       # tap = i*2+j
        y = acc * scale.view(1, 1, 1, -1, 1, 1) + shift.view(1, 1, 1, -1, 1, 1)
        y = y.permute(0, 4, 1, 5, 2, 3).reshape(n, 2 * oh, 2 * ow, cout)   # [n, 2h+i, 2w+j, c]
    else:
        y = acc * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        y = y.permute(0, 2, 3, 1)
        if res is not None:
            r = res.float()
            if res_up2:
                r = r.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
            y = y + r
    if relu:
        y = torch.relu(y)
    ld = out_ld or cout
    odt = torch.float32 if out_f32 else torch.bfloat16
    oshape = (n, 2 * oh, 2 * ow, ld) if deconv else (n, oh, ow, ld)
    desc = dict(n=n, h=h, w=w, cin=cin, kh=k, kw=k, stride=stride, pad=pad, cout=cout, relu=int(relu),
                residual_upsample2=int(res_up2), out_dtype=0 if out_f32 else 1, out_mode=int(deconv), out_ld=out_ld)
    w2 = wk.reshape(taps * cout, k * k * cin).contiguous()
    return x, w2, scale, shift, res, desc, oshape, odt, y.contiguous()


CASES = {
    "1x1_single_kblock": dict(n=2, h=16, w=16, cin=64, cout=64, k=1),
    "1x1_relu_residual": dict(n=2, h=16, w=16, cin=256, cout=256, k=1, relu=True, residual=True),
    "1x1_m_tail": dict(n=1, h=10, w=20, cin=128, cout=64, k=1, relu=True),
    "1x1_stride2": dict(n=2, h=32, w=32, cin=256, cout=512, k=1, stride=2, relu=True),
    "1x1_lateral_upsampled_residual": dict(n=2, h=16, w=16, cin=512, cout=256, k=1, residual=True, res_up2=True),
    "3x3_16": dict(n=2, h=16, w=16, cin=64, cout=64, k=3, relu=True),
    "3x3_64": dict(n=1, h=64, w=64, cin=128, cout=128, k=3, relu=True),
    "3x3_8_two_images_per_tile": dict(n=3, h=8, w=8, cin=256, cout=256, k=3),
    "3x3_4_eight_images_per_tile": dict(n=9, h=4, w=4, cin=256, cout=512, k=3, relu=True),
    "3x3_mask_head_14": dict(n=5, h=14, w=14, cin=256, cout=256, k=3, relu=True),
    "deconv_2x2": dict(n=3, h=14, w=14, cin=256, cout=256, k=1, relu=True, deconv=True),
    "rpn_head_f32_padded": dict(n=2, h=16, w=16, cin=512, cout=18, k=1, out_f32=True, out_ld=32),
    "fc_12544": dict(n=1, h=1, w=300, cin=12544, cout=1024, k=1, relu=True),
    "mask_logits_f32": dict(n=2, h=28, w=28, cin=256, cout=4, k=1, out_f32=True, out_ld=8),
    # many tiles per persistent CTA (ring + TMEM double buffer wrap several times), 256-wide tiles
    "1x1_many_tiles_bn256": dict(n=16, h=64, w=64, cin=128, cout=256, k=1, relu=True, residual=True),
    "1x1_many_tiles_k64": dict(n=8, h=64, w=64, cin=64, cout=64, k=1, relu=True),
    "3x3_mask_head_many_bn256": dict(n=200, h=14, w=14, cin=256, cout=256, k=3, relu=True),
    "deconv_many_bn256": dict(n=100, h=14, w=14, cin=256, cout=256, k=1, relu=True, deconv=True),
    "3x3_p2_like": dict(n=4, h=64, w=64, cin=256, cout=512, k=3, relu=True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_conv_tcgen05_matches_fp32_reference(name):
    x, w2, scale, shift, res, desc, oshape, odt, ref = make_case(hash(name) % 1000, **CASES[name])
    out = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    cout = desc["cout"]
    got = out[..., :cout].float()
    assert torch.isfinite(got).all(), "non-finite output (unwritten rows?)"
    err = (got - ref).abs()
    # bf16 store: |err| <= 2^-8 |ref| + accumulation noise; f32 store: accumulation noise only
    tol = (2.0 ** -7) * ref.abs() + 2e-2 if odt == torch.bfloat16 else 1e-3 * (1 + ref.abs())
    bad = (err > tol).sum().item()
    assert bad == 0, "%d elements out of tolerance, max err %g" % (bad, err.max().item())
    if oshape[-1] > cout:   # padded columns are written as zeros
        assert (out[..., cout:oshape[-1]].float() == 0).all()


@pytest.mark.parametrize("name", ["1x1_relu_residual", "3x3_16", "1x1_stride2", "deconv_2x2", "3x3_mask_head_14"])
def test_conv_tcgen05_matches_cuda_core_reference(name):
    x, w2, scale, shift, res, desc, oshape, odt, ref = make_case(7, **CASES[name])
    a = run_conv(x, w2, scale, shift, res, desc, oshape, odt).float()
    b = run_conv(x, w2, scale, shift, res, desc, oshape, odt, simt=True).float()
    # same operands, same epilogue, fp32 accumulation in a different order: <= 1 bf16 ulp
    assert ((a - b).abs() <= (2.0 ** -7) * b.abs() + 1e-3).all()


@pytest.mark.parametrize("name", [n for n in CASES if CASES[n]["k"] == 3])
def test_conv_im2col_mode_every_3x3_shape(name, monkeypatch):
    """TMA im2col-mode A loads forced on for every 3x3 shape (tile mode is the default where the map tiles
    into full 128-pixel rectangles): same result as tile mode to the last bit (same K order per element)."""
    x, w2, scale, shift, res, desc, oshape, odt, ref = make_case(11, **CASES[name])
    monkeypatch.setenv("MRCNN_B200_IM2COL", "0")
    a = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    monkeypatch.setenv("MRCNN_B200_IM2COL", "1")
    b = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    assert torch.isfinite(b.float()).all()
    assert torch.equal(a, b)
    err = (b.float() - ref).abs()
    assert (err <= (2.0 ** -7) * ref.abs() + 2e-2).all()


EPI_TMA_CASES = [n for n, c in CASES.items() if not c.get("deconv") and not c.get("out_f32") and not c.get("res_up2")
                 and c.get("stride", 1) == 1 and c["cout"] % 8 == 0]


@pytest.mark.parametrize("name", EPI_TMA_CASES)
def test_conv_tma_epilogue_is_bit_identical(name, monkeypatch):
    """Shared-memory epilogue (TMA residual loads + TMA output stores, 64B swizzle) against the direct-store
    epilogue: the arithmetic per element is the same, so the bf16 outputs must be equal bit for bit."""
    x, w2, scale, shift, res, desc, oshape, odt, ref = make_case(13, **CASES[name])
    monkeypatch.setenv("MRCNN_B200_EPI_TMA", "0")
    a = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    monkeypatch.setenv("MRCNN_B200_EPI_TMA", "1")
    b = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    assert torch.isfinite(b.float()).all(), "non-finite output (unwritten rows?)"
    assert torch.equal(a, b)
    err = (b.float() - ref).abs()
    assert (err <= (2.0 ** -7) * ref.abs() + 2e-2).all()


@pytest.mark.parametrize("epi", ["0", "1"])
@pytest.mark.parametrize("name", [n for n in CASES if n != "fc_12544"])
def test_conv_two_ctas_per_sm_variant_is_bit_identical(name, epi, monkeypatch):
    """OCC2 launch variant (two persistent CTAs per SM, shorter ring, two staging buffers per epilogue warp; the
    engine's autotuner picks it per layer) against the one-CTA-per-SM kernel: same K order per element, so the outputs
    must be equal bit for bit, with both epilogues."""
    x, w2, scale, shift, res, desc, oshape, odt, ref = make_case(17, **CASES[name])
    monkeypatch.setenv("MRCNN_B200_EPI_TMA", epi)
    monkeypatch.setenv("MRCNN_B200_OCC2", "0")
    a = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    monkeypatch.setenv("MRCNN_B200_OCC2", "2")
    b = run_conv(x, w2, scale, shift, res, desc, oshape, odt)
    cout = desc["cout"]
    assert torch.isfinite(b[..., :cout].float()).all(), "non-finite output (unwritten rows?)"
    assert torch.equal(a, b)


def test_conv_rejects_unsupported():
    nat = _native()
    lib = nat.lib()
    x = torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device="cuda")
    d = nat.ConvDesc(n=1, h=8, w=8, cin=48, kh=1, kw=1, stride=1, pad=0, cout=64, relu=0, residual_upsample2=0,
                     out_dtype=1, out_mode=0, out_ld=0)
    st = lib.mrcnn_conv2d_bf16(d, nat.ptr(x), nat.ptr(x), nat.ptr(x), nat.ptr(x), None, nat.ptr(x), None)
    assert st != 0 and b"multiple of 64" in lib.mrcnn_last_error()
