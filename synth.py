"""Synthetic FITS-like radio maps for the benchmark and the tests (SURVEY.md §8d
config #2): per image i, numpy.random.default_rng(1234 + i); S x S float32; Gaussian noise
sigma = 3e-4 Jy; 20-60 point sources with log-uniform peak 1e-3..5e-2 Jy convolved with an
elliptical Gaussian beam (FWHM 3-6 px); 0-3 extended double-lobe sources; an 8-px NaN strip on 25 %
of the images (exercises the NaN -> min fill)."""
import zlib

import numpy as np


def _gauss2d(S, y0, x0, fwhm_y, fwhm_x, theta, peak, yy, xx):
    sy, sx = fwhm_y / 2.3548, fwhm_x / 2.3548
    ct, st = np.cos(theta), np.sin(theta)
    dy, dx = yy - y0, xx - x0
    u = ct * dx + st * dy
    v = -st * dx + ct * dy
    return peak * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2))


def radio_map(i, S=256):
    rng = np.random.default_rng(1234 + i)
    img = rng.normal(0.0, 3e-4, size=(S, S))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float64)
    bmaj, bmin, bpa = rng.uniform(3, 6), rng.uniform(3, 6), rng.uniform(0, np.pi)
    for _ in range(int(rng.integers(20, 61))):
        peak = 10 ** rng.uniform(-3, np.log10(5e-2))
        y0, x0 = rng.uniform(0, S, 2)
        r = int(4 * max(bmaj, bmin))
        ys, ye = max(0, int(y0) - r), min(S, int(y0) + r + 1)
        xs, xe = max(0, int(x0) - r), min(S, int(x0) + r + 1)
        img[ys:ye, xs:xe] += _gauss2d(S, y0, x0, bmaj, bmin, bpa, peak, yy[ys:ye, xs:xe], xx[ys:ye, xs:xe])
    for _ in range(int(rng.integers(0, 4))):
        y0, x0 = rng.uniform(0.15 * S, 0.85 * S, 2)
        sep, ang = rng.uniform(6, 0.12 * S), rng.uniform(0, np.pi)
        peak = 10 ** rng.uniform(-2.5, -1.5)
        for sgn in (-1, 1):
            cy, cx = y0 + sgn * sep * np.sin(ang), x0 + sgn * sep * np.cos(ang)
            img += _gauss2d(S, cy, cx, rng.uniform(6, 14), rng.uniform(4, 8), ang, peak, yy, xx)
    img = img.astype(np.float32)
    if rng.random() < 0.25:
        k = int(rng.integers(0, 4))
        if k == 0:
            img[:8, :] = np.nan
        elif k == 1:
            img[-8:, :] = np.nan
        elif k == 2:
            img[:, :8] = np.nan
        else:
            img[:, -8:] = np.nan
    return img


def radio_maps(n, S=256, start=0):
    return np.stack([radio_map(start + i, S) for i in range(n)])


# --------------------------------------------------------------------------------------------
# weighted-layer inventory of the inference graph + seeded random weights
# --------------------------------------------------------------------------------------------

def layer_specs(num_classes=4, fc_size=1024, pyramid=256):
    """Ordered list of (layer_name, kind, kernel_shape) with kind in conv|bn|dense|deconv."""
    specs = [("conv1", "conv", (7, 7, 3, 64)), ("bn_conv1", "bn", (64,))]
    cin = 64
    stages = [(2, "abc", (64, 64, 256)), (3, "abcd", (128, 128, 512)),
              (4, "a" + "".join(chr(98 + i) for i in range(22)), (256, 256, 1024)),
              (5, "abc", (512, 512, 2048))]
    for stage, blocks, (f1, f2, f3) in stages:
        for blk in blocks:
            base = "res%d%s_branch" % (stage, blk)
            bnb = "bn%d%s_branch" % (stage, blk)
            specs += [(base + "2a", "conv", (1, 1, cin, f1)), (bnb + "2a", "bn", (f1,)),
                      (base + "2b", "conv", (3, 3, f1, f2)), (bnb + "2b", "bn", (f2,)),
                      (base + "2c", "conv", (1, 1, f2, f3)), (bnb + "2c", "bn", (f3,))]
            if blk == "a":
                specs += [(base + "1", "conv", (1, 1, cin, f3)), (bnb + "1", "bn", (f3,))]
            cin = f3
    for name, c in (("fpn_c5p5", 2048), ("fpn_c4p4", 1024), ("fpn_c3p3", 512), ("fpn_c2p2", 256)):
        specs.append((name, "conv", (1, 1, c, pyramid)))
    for name in ("fpn_p2", "fpn_p3", "fpn_p4", "fpn_p5"):
        specs.append((name, "conv", (3, 3, pyramid, pyramid)))
    specs += [("rpn_conv_shared", "conv", (3, 3, pyramid, 512)),
              ("rpn_class_raw", "conv", (1, 1, 512, 6)),
              ("rpn_bbox_pred", "conv", (1, 1, 512, 12))]
    specs += [("mrcnn_class_conv1", "conv", (7, 7, pyramid, fc_size)),
              ("mrcnn_class_bn1", "bn", (fc_size,)),
              ("mrcnn_class_conv2", "conv", (1, 1, fc_size, fc_size)),
              ("mrcnn_class_bn2", "bn", (fc_size,)),
              ("mrcnn_class_logits", "dense", (fc_size, num_classes)),
              ("mrcnn_bbox_fc", "dense", (fc_size, 4 * num_classes))]
    for i in range(1, 5):
        specs += [("mrcnn_mask_conv%d" % i, "conv", (3, 3, pyramid, pyramid)),
                  ("mrcnn_mask_bn%d" % i, "bn", (pyramid,))]
    specs += [("mrcnn_mask_deconv", "deconv", (2, 2, pyramid, pyramid)),
              ("mrcnn_mask", "conv", (1, 1, pyramid, num_classes))]
    return specs


# per-layer gain on the He-normal std, tuned so that with inputs in 0..255 the pyramid has
# std ~1-2, RPN/class logits std ~2 and box deltas std ~1 (a non-degenerate detect workload)
_GAIN = {"conv1": 1.0 / 64.0, "fpn_c5p5": 0.12, "fpn_c4p4": 0.12, "fpn_c3p3": 0.2, "fpn_c2p2": 0.2,
         "fpn_p": 0.6, "rpn_conv_shared": 0.7, "rpn_class_raw": 1.6, "rpn_bbox_pred": 0.8,
         "mrcnn_class_logits": 1.5, "mrcnn_bbox_fc": 0.8, "mrcnn_mask": 1.5}


def make_random_weights(seed=0, num_classes=4):
    """Deterministic stand-in for share/mrcnn_weights.h5 (an unresolved LFS pointer).
    Returns {layer_name: [arrays in Keras layer.weights order]} (float32)."""
    out = {}
    for name, kind, shape in layer_specs(num_classes):
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if kind == "bn":
            c = shape[0]
            gamma = rng.uniform(0.6, 1.0, c)
            if name.endswith("_branch2c"):
                gamma = rng.uniform(0.2, 0.4, c)      # keep the residual trunk from blowing up
            beta = rng.normal(0.0, 0.1, c)
            mean = rng.normal(0.0, 0.1, c)
            var = rng.uniform(0.5, 1.5, c)
            out[name] = [a.astype(np.float32) for a in (gamma, beta, mean, var)]
        else:
            if kind == "dense":
                fan_in = shape[0]
                cout = shape[1]
            elif kind == "deconv":
                fan_in = shape[3]
                cout = shape[2]
            else:
                fan_in = shape[0] * shape[1] * shape[2]
                cout = shape[3]
            std = np.sqrt(2.0 / fan_in) * _GAIN.get(name, _GAIN.get(name.rstrip("0123456789"), 1.0))
            k = rng.normal(0.0, std, shape)
            b = rng.normal(0.0, 0.05, cout)
            out[name] = [k.astype(np.float32), b.astype(np.float32)]
    return out


# --------------------------------------------------------------------------------------------
# training samples (BASELINE.json configs[4]: synthetic GT from the generator's own source list)
# --------------------------------------------------------------------------------------------

def training_sample(i, S=256):
    """-> (map [S,S] float32, masks [S,S,n] bool, class_ids [n] int32).  Sources are drawn like radio_map's (compact
    sources = class 2 'source', double-lobe sources = class 3 'galaxy', a faint ring artefact around the brightest compact
    source = class 1 'sidelobe'); the mask of a source is where its own flux exceeds 3 sigma of the noise."""
    rng = np.random.default_rng(99000 + i)
    sigma = 3e-4
    img = rng.normal(0.0, sigma, size=(S, S))
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float64)
    bmaj, bmin, bpa = rng.uniform(3, 6), rng.uniform(3, 6), rng.uniform(0, np.pi)
    masks, cls, peaks = [], [], []
    for _ in range(int(rng.integers(6, 20))):
        peak = 10 ** rng.uniform(-2.6, np.log10(5e-2))
        y0, x0 = rng.uniform(8, S - 8, 2)
        src = _gauss2d(S, y0, x0, bmaj, bmin, bpa, peak, yy, xx)
        img += src
        masks.append(src > 3 * sigma)
        cls.append(2)
        peaks.append((peak, y0, x0))
    for _ in range(int(rng.integers(1, 4))):
        y0, x0 = rng.uniform(0.2 * S, 0.8 * S, 2)
        sep, ang = rng.uniform(6, 0.1 * S), rng.uniform(0, np.pi)
        peak = 10 ** rng.uniform(-2.5, -1.5)
        src = np.zeros((S, S))
        for sgn in (-1, 1):
            src += _gauss2d(S, y0 + sgn * sep * np.sin(ang), x0 + sgn * sep * np.cos(ang), rng.uniform(6, 12), rng.uniform(4, 8),
                            ang, peak, yy, xx)
        img += src
        masks.append(src > 3 * sigma)
        cls.append(3)
    peak, y0, x0 = max(peaks)
    rr = np.hypot(yy - y0, xx - x0)
    ring = 0.02 * peak * np.exp(-0.5 * ((rr - 4 * max(bmaj, bmin)) / 1.5) ** 2)
    img += ring
    if (ring > 2 * sigma).any():
        masks.append(ring > 2 * sigma)
        cls.append(1)
    masks = np.stack(masks, axis=-1)
    keep = masks.sum(axis=(0, 1)) > 0
    return img.astype(np.float32), masks[:, :, keep], np.asarray(cls, dtype=np.int32)[keep]
