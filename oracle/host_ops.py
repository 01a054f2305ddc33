"""ORACLE (test infrastructure) — numpy restatement of the host-side pieces of the detect path.

  zscale_limits / zscale_stretch  <- astropy.visualization.ZScaleInterval (third-party, absent;
                                     call site mrcnn/utils.py:1166-1172)
  fits_to_rgb                     <- mrcnn/utils.py:1081-1163 (read_fits after the FITS decode),
                                     normalize_img :1182-1188, gray2rgb :1190-1208
  skimage_resize                  <- skimage.transform.resize <=0.15 (third-party, absent; wrapper
                                     mrcnn/utils.py:957-978)
  resize_image                    <- mrcnn/utils.py:456-561 ("square" / "none" / "pad64")
  mold_image / compose_image_meta / mold_inputs <- mrcnn/model.py:2964-2969, 2891-2913, 2519-2556
  generate_anchors / generate_pyramid_anchors / norm_boxes / denorm_boxes
                                  <- mrcnn/utils.py:652-708, 923-954
  compute_backbone_shapes         <- mrcnn/model.py:75-89
  unmold_mask / unmold_detections <- mrcnn/utils.py:629-645, mrcnn/model.py:2558-2621

parity: zscale and skimage_resize are UNPINNED (third-party algorithm restated); the pure-numpy
pieces are pinned by tests/golden/ref_numpy_golden.npz (made from the real reference functions).
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# a1: FITS -> uint8 RGB
# --------------------------------------------------------------------------------------------

def parse_fits_primary(raw):
    """Minimal FITS primary-HDU decoder (2-D or 4-D image, BITPIX -32/-64/16/32).
    Independent twin of the product reader; returns (data ndarray, header dict)."""
    header = {}
    pos = 0
    done = False
    while not done:
        block = raw[pos:pos + 2880]
        if len(block) < 2880:
            raise ValueError("truncated FITS header")
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                s = val.lstrip()[1:]
                end = s.find("'")
                header[key] = s[:end].rstrip()
            else:
                v = val.split("/")[0].strip()
                if v in ("T", "F"):
                    header[key] = (v == "T")
                else:
                    try:
                        header[key] = int(v)
                    except ValueError:
                        try:
                            header[key] = float(v.replace("D", "E"))
                        except ValueError:
                            header[key] = v
        pos += 2880
    bitpix = header["BITPIX"]
    naxis = header["NAXIS"]
    shape = [header["NAXIS%d" % (i + 1)] for i in range(naxis)][::-1]
    dt = {-32: ">f4", -64: ">f8", 16: ">i2", 32: ">i4", 8: "u1"}[bitpix]
    count = int(np.prod(shape)) if shape else 0
    data = np.frombuffer(raw, dtype=dt, count=count, offset=pos).reshape(shape)
    bscale = header.get("BSCALE", 1.0)
    bzero = header.get("BZERO", 0.0)
    if bitpix > 0 and (bscale != 1.0 or bzero != 0.0):
        data = data * bscale + bzero
    return data, header


def zscale_limits(values, contrast=0.25, nsamples=1000, max_reject=0.5, min_npixels=5,
                  krej=2.5, max_iterations=5):
    """astropy ZScaleInterval.get_limits. Returns (vmin, vmax) as the numpy scalars astropy
    would return (float32 sample or float64 expression)."""
    values = np.asarray(values)
    values = values[np.isfinite(values)]
    stride = int(max(1.0, values.size / nsamples))
    samples = values[::stride][:nsamples].copy()
    samples.sort()
    npix = len(samples)
    vmin = samples[0]
    vmax = samples[-1]
    minpix = max(min_npixels, int(npix * max_reject))
    x = np.arange(npix)
    ngoodpix = npix
    last_ngoodpix = npix + 1
    badpix = np.zeros(npix, dtype=bool)
    ngrow = max(1, int(npix * 0.01))
    kernel = np.ones(ngrow, dtype=bool)
    fit = None
    for _ in range(max_iterations):
        if ngoodpix >= last_ngoodpix or ngoodpix < minpix:
            break
        fit = np.polyfit(x, samples, deg=1, w=(~badpix).astype(int))
        fitted = np.poly1d(fit)(x)
        flat = samples - fitted
        threshold = krej * flat[~badpix].std()
        badpix[(flat < -threshold) | (flat > threshold)] = True
        badpix = np.convolve(badpix, kernel, mode="same")
        last_ngoodpix = ngoodpix
        ngoodpix = np.sum(~badpix)
    slope, _intercept = fit
    if ngoodpix >= minpix:
        if contrast > 0:
            slope = slope / contrast
        center_pixel = (npix - 1) // 2
        median = np.median(samples)
        vmin = max(vmin, median - (center_pixel - 1) * slope)
        vmax = min(vmax, median + (npix - center_pixel) * slope)
    return vmin, vmax


def zscale_apply(data, vmin, vmax):
    """ZScaleInterval.__call__ after the limits: float32 array in -> float32 out, with the
    legacy (numpy < 2, as pinned by the reference era) scalar casting: both the offset and the
    divisor are rounded to float32 before use."""
    data = np.asarray(data, dtype=F32)
    out = data - F32(float(vmin))
    rng = vmax - vmin                                   # float64 when either side is float64
    if rng != 0:
        out = out / F32(rng)
    return np.clip(out, F32(0.0), F32(1.0)).astype(F32)


def zscale_stretch(data, contrast=0.25):
    vmin, vmax = zscale_limits(data, contrast)
    return zscale_apply(data, vmin, vmax)


def fits_to_rgb(data, zscale_contrasts=(0.25, 0.25, 0.25), stretch=True, normalize=True,
                to_uint8=True):
    """mrcnn/utils.py:1081-1163 from the float conversion on. data: 2-D array."""
    x = np.array(data, dtype=F32)                       # :1081
    img_min = np.nanmin(x)                              # :1090
    x[np.isnan(x)] = img_min                            # :1091
    chans = []
    for c in range(3):
        ch = x.copy()
        if stretch:
            ch = zscale_stretch(ch, zscale_contrasts[c]).astype(F32)   # :1101-1111
        if normalize:
            ch = (ch / np.max(ch)).astype(F32)          # :1182-1188
        chans.append(ch)
    if to_uint8:
        out = [np.array((ch * F32(255)).round(), dtype=np.uint8) for ch in chans]   # :1196-1198
    else:
        out = [np.array(ch * F32(255), dtype=F32) for ch in chans]                  # :1200-1202
    return np.stack(out, axis=-1)


# --------------------------------------------------------------------------------------------
# a2: skimage<=0.15 resize, resize_image, mold
# --------------------------------------------------------------------------------------------

def skimage_resize(image, output_shape, preserve_range=False):
    """skimage.transform.resize(order=1, mode='constant', cval=0, clip=True,
    anti_aliasing=False) as scikit-image <= 0.15 computes it (affine warp, float64):

      c = col_scale*x + (0.5*col_scale - 0.5),  r = row_scale*y + (0.5*row_scale - 0.5)
      out = (1-dr)*((1-dc)*p[r0,c0] + dc*p[r0,c1]) + dr*((1-dc)*p[r1,c0] + dc*p[r1,c1])
      with p = cval (0) for any neighbour outside the image, r0/c0 = floor, r1/c1 = ceil;
      finally clip to [image.min(), image.max()] (pixels equal to cval are left alone when cval
      lies outside that range).
    Integer inputs with preserve_range=False are scaled like img_as_float (uint8 -> /255,
    bool -> 0/1).
    """
    image = np.asarray(image)
    if preserve_range or image.dtype.kind == "f":
        img = image.astype(np.float64)
    elif image.dtype == np.bool_:
        img = image.astype(np.float64)
    elif image.dtype == np.uint8:
        img = image.astype(np.float64) / 255.0
    else:
        raise NotImplementedError(image.dtype)
    rows, cols = int(output_shape[0]), int(output_shape[1])
    in_rows, in_cols = img.shape[0], img.shape[1]
    squeeze = (img.ndim == 2)
    if squeeze:
        img = img[:, :, None]
    if rows == 0 or cols == 0:
        res = np.zeros((rows, cols, img.shape[2]), dtype=np.float64)
        return res[:, :, 0] if squeeze else res
    row_scale = np.float64(in_rows) / np.float64(rows)
    col_scale = np.float64(in_cols) / np.float64(cols)
    if rows == 1 and cols == 1:
        # translation-only transform of skimage for a 1x1 output
        r = np.array([in_rows / 2.0 - 0.5])
        c = np.array([in_cols / 2.0 - 0.5])
    else:
        r = row_scale * np.arange(rows, dtype=np.float64) + (0.5 * row_scale - 0.5)
        c = col_scale * np.arange(cols, dtype=np.float64) + (0.5 * col_scale - 0.5)
    r0 = np.floor(r).astype(np.int64)
    r1 = np.ceil(r).astype(np.int64)
    c0 = np.floor(c).astype(np.int64)
    c1 = np.ceil(c).astype(np.int64)
    dr = (r - r0)[:, None, None]
    dc = (c - c0)[None, :, None]

    def px(ri, ci):
        vr = (ri >= 0) & (ri < in_rows)
        vc = (ci >= 0) & (ci < in_cols)
        v = img[np.clip(ri, 0, in_rows - 1)[:, None], np.clip(ci, 0, in_cols - 1)[None, :]]
        return np.where((vr[:, None] & vc[None, :])[:, :, None], v, 0.0)

    top = (1 - dc) * px(r0, c0) + dc * px(r0, c1)
    bottom = (1 - dc) * px(r1, c0) + dc * px(r1, c1)
    out = (1 - dr) * top + dr * bottom
    mn, mx = img.min(), img.max()
    preserve_cval = not (mn <= 0.0 <= mx)
    if preserve_cval:
        cmask = (out == 0.0)
    out = np.clip(out, mn, mx)
    if preserve_cval:
        out[cmask] = 0.0
    return out[:, :, 0] if squeeze else out


def resize_image(image, min_dim=None, max_dim=None, min_scale=None, mode="square"):
    """mrcnn/utils.py:456-561 (modes none/square/pad64; 'crop' is training-only)."""
    image_dtype = image.dtype
    h, w = image.shape[:2]
    window = (0, 0, h, w)
    scale = 1
    padding = [(0, 0), (0, 0), (0, 0)]
    crop = None
    if mode == "none":
        return image, window, scale, padding, crop
    if min_dim:
        scale = max(1, min_dim / min(h, w))
    if min_scale and scale < min_scale:
        scale = min_scale
    if max_dim and mode == "square":
        image_max = max(h, w)
        if round(image_max * scale) > max_dim:
            scale = max_dim / image_max
    if scale != 1:
        image = skimage_resize(image, (round(h * scale), round(w * scale)), preserve_range=True)
    if mode == "square":
        h, w = image.shape[:2]
        top_pad = (max_dim - h) // 2
        bottom_pad = max_dim - h - top_pad
        left_pad = (max_dim - w) // 2
        right_pad = max_dim - w - left_pad
        padding = [(top_pad, bottom_pad), (left_pad, right_pad), (0, 0)]
        image = np.pad(image, padding, mode="constant", constant_values=0)
        window = (top_pad, left_pad, h + top_pad, w + left_pad)
    elif mode == "pad64":
        h, w = image.shape[:2]
        assert min_dim % 64 == 0
        if h % 64 > 0:
            max_h = h - (h % 64) + 64
            top_pad = (max_h - h) // 2
            bottom_pad = max_h - h - top_pad
        else:
            top_pad = bottom_pad = 0
        if w % 64 > 0:
            max_w = w - (w % 64) + 64
            left_pad = (max_w - w) // 2
            right_pad = max_w - w - left_pad
        else:
            left_pad = right_pad = 0
        padding = [(top_pad, bottom_pad), (left_pad, right_pad), (0, 0)]
        image = np.pad(image, padding, mode="constant", constant_values=0)
        window = (top_pad, left_pad, h + top_pad, w + left_pad)
    else:
        raise Exception("Mode {} not supported".format(mode))
    return image.astype(image_dtype), window, scale, padding, crop


def mold_image(images, mean_pixel):
    return images.astype(F32) - np.asarray(mean_pixel)


def compose_image_meta(image_id, original_image_shape, image_shape, window, scale,
                       active_class_ids):
    return np.array([image_id] + list(original_image_shape) + list(image_shape) + list(window)
                    + [scale] + list(active_class_ids))


def mold_inputs(images, *, min_dim, max_dim, min_scale, mode, mean_pixel, num_classes):
    """mrcnn/model.py:2519-2556 -> (molded [B,S,S,3] float32 as fed to TF, metas, windows)."""
    molded, metas, windows = [], [], []
    for image in images:
        m, window, scale, _pad, _crop = resize_image(image, min_dim=min_dim, min_scale=min_scale,
                                                     max_dim=max_dim, mode=mode)
        m = mold_image(m, mean_pixel)
        meta = compose_image_meta(0, image.shape, m.shape, window, scale,
                                  np.zeros([num_classes], dtype=np.int32))
        molded.append(m)
        windows.append(window)
        metas.append(meta)
    return (np.stack(molded).astype(F32), np.stack(metas), np.stack(windows))


# --------------------------------------------------------------------------------------------
# a3: anchors
# --------------------------------------------------------------------------------------------

def compute_backbone_shapes(image_shape, strides=(4, 8, 16, 32, 64)):
    return np.array([[int(math.ceil(image_shape[0] / s)), int(math.ceil(image_shape[1] / s))]
                     for s in strides])


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    scales, ratios = np.meshgrid(np.array(scales), np.array(ratios))
    scales = scales.flatten()
    ratios = ratios.flatten()
    heights = scales / np.sqrt(ratios)
    widths = scales * np.sqrt(ratios)
    shifts_y = np.arange(0, shape[0], anchor_stride) * feature_stride
    shifts_x = np.arange(0, shape[1], anchor_stride) * feature_stride
    shifts_x, shifts_y = np.meshgrid(shifts_x, shifts_y)
    box_widths, box_centers_x = np.meshgrid(widths, shifts_x)
    box_heights, box_centers_y = np.meshgrid(heights, shifts_y)
    box_centers = np.stack([box_centers_y, box_centers_x], axis=2).reshape([-1, 2])
    box_sizes = np.stack([box_heights, box_widths], axis=2).reshape([-1, 2])
    return np.concatenate([box_centers - 0.5 * box_sizes, box_centers + 0.5 * box_sizes], axis=1)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i],
                                            feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


def norm_boxes(boxes, shape):
    h, w = shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.divide((boxes - shift), scale).astype(F32)


def denorm_boxes(boxes, shape):
    h, w = shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.around(np.multiply(boxes, scale) + shift).astype(np.int32)


def get_anchors(image_shape, scales, ratios=(0.5, 1, 2), strides=(4, 8, 16, 32, 64),
                anchor_stride=1):
    """mrcnn/model.py:2764-2784 -> [A,4] float32 normalized."""
    shapes = compute_backbone_shapes(image_shape, strides)
    a = generate_pyramid_anchors(scales, ratios, shapes, strides, anchor_stride)
    return norm_boxes(a, image_shape[:2])


# --------------------------------------------------------------------------------------------
# a12: unmold
# --------------------------------------------------------------------------------------------

def unmold_mask(mask, bbox, image_shape):
    """mrcnn/utils.py:629-645."""
    y1, x1, y2, x2 = [int(v) for v in bbox]
    m = skimage_resize(mask, (y2 - y1, x2 - x1))
    m = np.where(m >= 0.5, 1, 0).astype(bool)
    full = np.zeros(image_shape[:2], dtype=bool)
    full[y1:y2, x1:x2] = m
    return full


def unmold_detections(detections, mrcnn_mask, original_image_shape, image_shape, window):
    """mrcnn/model.py:2558-2621 for one image."""
    detections = np.asarray(detections)
    zero_ix = np.where(detections[:, 4] == 0)[0]
    N = zero_ix[0] if zero_ix.shape[0] > 0 else detections.shape[0]
    boxes = detections[:N, :4]
    class_ids = detections[:N, 4].astype(np.int32)
    scores = detections[:N, 5]
    masks = mrcnn_mask[np.arange(N), :, :, class_ids]
    window = norm_boxes(np.asarray(window), image_shape[:2])
    wy1, wx1, wy2, wx2 = window
    shift = np.array([wy1, wx1, wy1, wx1])
    wh = wy2 - wy1
    ww = wx2 - wx1
    scale = np.array([wh, ww, wh, ww])
    boxes = np.divide(boxes - shift, scale)
    boxes = denorm_boxes(boxes, original_image_shape[:2])
    exclude_ix = np.where((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) <= 0)[0]
    if exclude_ix.shape[0] > 0:
        boxes = np.delete(boxes, exclude_ix, axis=0)
        class_ids = np.delete(class_ids, exclude_ix, axis=0)
        scores = np.delete(scores, exclude_ix, axis=0)
        masks = np.delete(masks, exclude_ix, axis=0)
        N = class_ids.shape[0]
    full_masks = [unmold_mask(masks[i], boxes[i], original_image_shape) for i in range(N)]
    full_masks = np.stack(full_masks, axis=-1) if full_masks \
        else np.empty(tuple(original_image_shape[:2]) + (0,))
    return boxes, class_ids, scores, full_masks
