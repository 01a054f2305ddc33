"""ctypes binding of libmrcnn_b200.so (C ABI: include/mrcnn_b200.h).

There is no CPU fallback: if the library is missing, `lib()` raises with the build command, and
every compute entry point needs a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libmrcnn_b200.so")

DTYPE_F32 = 0
DTYPE_BF16 = 1

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t


class NativeError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    _fields_ = [("n", c_int), ("h", c_int), ("w", c_int), ("cin", c_int),
                ("kh", c_int), ("kw", c_int), ("stride", c_int), ("pad", c_int),
                ("cout", c_int), ("relu", c_int), ("residual_upsample2", c_int),
                ("out_dtype", c_int), ("out_mode", c_int), ("out_ld", c_int)]


class EngineConfig(ctypes.Structure):
    _fields_ = [("batch_size", c_int), ("image_size", c_int), ("num_classes", c_int),
                ("pre_nms_limit", c_int), ("post_nms_rois", c_int),
                ("detection_max_instances", c_int), ("pool_size", c_int),
                ("mask_pool_size", c_int), ("fc_layers_size", c_int),
                ("top_down_pyramid_size", c_int), ("anchors_per_location", c_int),
                ("rpn_nms_threshold", c_float), ("detection_min_confidence", c_float),
                ("detection_nms_threshold", c_float), ("rpn_bbox_std_dev", c_float * 4),
                ("bbox_std_dev", c_float * 4), ("backbone_strides", c_int * 5)]


# name -> (restype, argtypes); every symbol include/mrcnn_b200.h declares
SIGNATURES = {
    "mrcnn_last_error": (ctypes.c_char_p, []),
    "mrcnn_abi_version": (c_int, []),
    "mrcnn_kernel_launch_count": (ctypes.c_ulonglong, []),
    "mrcnn_zscale_params": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mrcnn_stretch_to_rgb8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mrcnn_resize_pad_mold": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p]),
    "mrcnn_skimage_resize_f64": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_double, ctypes.c_double,
                                         c_void_p, c_void_p]),
    "mrcnn_proposal_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mrcnn_proposal_layer": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mrcnn_roi_levels": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "mrcnn_pyramid_roi_align": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                        c_float, c_void_p, c_void_p, c_void_p]),
    "mrcnn_detection_layer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "mrcnn_unmold_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mrcnn_unmold_detections": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "mrcnn_mask_bits_words": (c_int, [c_int]),
    "mrcnn_unmold_detections_bits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_size_t, c_void_p]),
    "mrcnn_host_threads": (c_int, []),
    "mrcnn_host_expand_mask_bits": (c_int, [c_void_p, c_int, ctypes.c_int64, c_int, c_void_p, c_void_p, c_int]),
    "mrcnn_conv2d_bf16": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "mrcnn_conv2d_bf16_simt": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "mrcnn_engine_create": (c_int, [ctypes.POINTER(EngineConfig), c_int, ctypes.POINTER(c_void_p)]),
    "mrcnn_engine_destroy": (None, [c_void_p]),
    "mrcnn_engine_num_layers": (c_int, [c_void_p]),
    "mrcnn_engine_layer_info": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_int),
                                        ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mrcnn_engine_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_int, c_void_p, c_size_t]),
    "mrcnn_engine_finalize": (c_int, [c_void_p, c_int]),
    "mrcnn_engine_set_anchors": (c_int, [c_void_p, c_void_p, c_int]),
    "mrcnn_engine_predict": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "mrcnn_engine_tensor": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_size_t)]),
    "mrcnn_engine_read": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, c_size_t]),
    "mrcnn_engine_run_stage": (c_int, [c_void_p, ctypes.c_char_p]),
    "mrcnn_engine_write": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, c_size_t]),
    "mrcnn_engine_detect_molded": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p]),
    "mrcnn_engine_detect_maps": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "mrcnn_engine_wait": (c_int, [c_void_p]),
    "mrcnn_engine_next_slot": (c_int, [c_void_p]),
    "mrcnn_engine_wait_slot": (c_int, [c_void_p, c_int]),
    "mrcnn_engine_stream": (c_void_p, [c_void_p]),
    "mrcnn_engine_stage_times": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_float)]),
    "mrcnn_engine_flops": (ctypes.c_double, [c_void_p]),
    "mrcnn_engine_set_profiling": (c_int, [c_void_p, c_int]),
    "mrcnn_engine_step_info": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                       ctypes.POINTER(c_float), ctypes.POINTER(ctypes.c_double)]),
    "mrcnn_engine_kernel_times": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_float),
                                          ctypes.POINTER(c_int)]),
    "mrcnn_mask_bits_expand_device": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_int64, c_int, c_void_p, c_void_p]),
    "mrcnn_engine_set_dense_output": (c_int, [c_void_p, c_void_p, c_int]),
    "mrcnn_engine_wait_slot_packed": (c_int, [c_void_p, c_int]),
    "mrcnn_engine_dense_copy_ms": (c_int, [c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_float)]),
    "mrcnn_plane_words": (c_size_t, [c_int, c_int]),
    "mrcnn_masks_pack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mrcnn_planes_area_bbox": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mrcnn_planes_pair_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mrcnn_planes_union": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mrcnn_planes_label_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mrcnn_planes_label": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mrcnn_labels_select": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mrcnn_planes_pixels": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mrcnn_planes_unpack": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mrcnn_pixel_lists_adjacent": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mrcnn_host_merge_components": (c_int, [c_int, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "mrcnn_host_all_pairs": (c_int, [c_int, c_void_p, c_void_p]),
    "mrcnn_host_pair_flags": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "mrcnn_host_contours": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mrcnn_host_contours_fetch": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mrcnn_conv2d_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mrcnn_conv2d_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mrcnn_shuffle_key": (ctypes.c_uint32, [ctypes.c_ulonglong, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]),
    "mrcnn_detection_targets": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_float, c_void_p, c_int, c_int, ctypes.c_ulonglong, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "mrcnn_pyramid_roi_align_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                                 c_void_p, c_void_p]),
    "mrcnn_conv_backward_prep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int,
                                         c_void_p]),
    "mrcnn_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_int, c_float,
                               c_float, c_float, c_float, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library (once). Raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError("libmrcnn_b200.so not built: run `python caesar-mrcnn_b200/build.py` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        partial = os.environ.get("MRCNN_B200_PARTIAL") == "1"     # development only
        for name, (res, args) in SIGNATURES.items():
            if partial and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)        # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().mrcnn_last_error()
        raise NativeError("%s failed (%d): %s" % (what or "native call", status, msg.decode() if msg else "?"))


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def float_array(values):
    return (c_float * len(values))(*[float(v) for v in values])
