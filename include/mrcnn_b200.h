/*
 * mrcnn_b200.h — C ABI of libmrcnn_b200.so: the B200 (sm_100a) implementation of the
 * caesar-mrcnn Mask R-CNN *detect* hot path.
 *
 * The reference (SKA-INAF/caesar-mrcnn) is pure Python on TensorFlow 1.13 / Keras 2.2 and has no
 * FFI of its own; its plug-in boundary for this path is the Python class surface
 * mrcnn.model.MaskRCNN / mrcnn.config.Config.  This header is the boundary one
 * level below it: one entry point per graph layer / host utility of the path, so that a
 * maintainer can bind each of them from the reference's own call sites with ctypes
 * (INTEGRATION.md shows the stubs), plus an engine object that runs the whole
 * keras_model.predict() of mode='inference'.  Each declaration cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  `stream` is a cudaStream_t passed as
 *     void* (NULL = default stream).  Unless stated otherwise pointers are DEVICE pointers and
 *     calls are asynchronous on `stream`.
 *   - tensors are dense, C-order, NHWC; boxes are (y1, x1, y2, x2).
 *   - every function returns 0 (MRCNN_STATUS_OK) or a negative status; mrcnn_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MRCNN_B200_H_
#define MRCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRCNN_STATUS_OK 0
#define MRCNN_STATUS_INVALID (-1)
#define MRCNN_STATUS_CUDA (-2)
#define MRCNN_STATUS_UNSUPPORTED (-3)
#define MRCNN_STATUS_NOTFOUND (-4)

#define MRCNN_DTYPE_F32 0
#define MRCNN_DTYPE_BF16 1

/* ---- library ------------------------------------------------------------------------------ */
const char* mrcnn_last_error(void);
int mrcnn_abi_version(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
unsigned long long mrcnn_kernel_launch_count(void);

/* ---- a1: FITS map -> uint8 RGB  (mrcnn/utils.py:1081-1208 read_fits after the FITS decode,
 *      stretch_img :1166-1172 = astropy ZScaleInterval, normalize_img :1182-1188, gray2rgb :1190-1208)
 * maps [n,H,W] float32 (NaN allowed).  params [n,3,4] float32 = per channel
 * (nan_fill, vmin, range, zmax): pixel -> clip((x-vmin)/range,0,1)/zmax.  One CTA per image. */
int mrcnn_zscale_params(const float* maps, int n_images, int height, int width,
                        const float* contrasts3, float* params, void* stream);
/* rgb [n,H,W,3] uint8 = round_half_even(255 * normalised stretch); minmax [n,2] int32 receives
 * the per-image min / max byte (needed by the skimage clip of the following resize). */
int mrcnn_stretch_to_rgb8(const float* maps, const float* params, int n_images, int height,
                          int width, uint8_t* rgb, int32_t* minmax, void* stream);

/* ---- a2: mold_inputs  (mrcnn/model.py:2519-2556; utils.resize_image mrcnn/utils.py:456-561 mode
 *      "square"; utils.resize :957-978 = skimage<=0.15 bilinear warp, cval 0, clip; mold_image
 *      mrcnn/model.py:2964-2969).  All n images share one original size (height,width).
 * rgb [n,H,W,3] uint8, minmax [n,2] (from mrcnn_stretch_to_rgb8) -> molded [n,S,S,3] float32:
 * resize to (out_h,out_w) (skipped when equal to (H,W)), truncate to uint8, paste at (top,left)
 * into a zero S x S frame, subtract mean_pixel (host pointer, 3 floats). */
int mrcnn_resize_pad_mold(const uint8_t* rgb, const int32_t* minmax, int n_images, int height,
                          int width, int out_h, int out_w, int square, int top, int left,
                          const float* mean_pixel3, float* molded, void* stream);

/* utils.resize (mrcnn/utils.py:957-978 = skimage.transform.resize(order=1, mode='constant', cval=0, clip=True,
 * anti_aliasing=False), scikit-image <= 0.15) for one float64 image [H,W,C] -> out [out_h,out_w,C] float64;
 * image_min / image_max = the clip range (min / max of the input).  Used by utils.resize / utils.unmold_mask. */
int mrcnn_skimage_resize_f64(const double* image, int height, int width, int channels, int out_h, int out_w,
                             double image_min, double image_max, double* out, void* stream);

/* ---- a7: ProposalLayer  (mrcnn/model.py:329-406, apply_box_deltas_graph :287-308,
 *      clip_boxes_graph :311-326, utils.batch_slice mrcnn/utils.py:872-906)
 * rpn_class [B,A,2], rpn_bbox [B,A,4], anchors [A,4] (anchors_batched=0) or [B,A,4] -> rpn_rois
 * [B,R,4] zero padded.  Optional taps: topk_idx [B,min(K,A)] (tf.nn.top_k order), keep_idx [B,R]
 * (NMS picks as indices into that order, -1 padded), keep_count [B].  If topk_idx is NULL a
 * workspace of mrcnn_proposal_workspace_bytes() must be given.  bbox_std_dev: HOST, 4 floats. */
size_t mrcnn_proposal_workspace_bytes(int batch, int num_anchors, int pre_nms_limit);
int mrcnn_proposal_layer(const float* rpn_class, const float* rpn_bbox, const float* anchors,
                         int anchors_batched, int batch, int num_anchors, int pre_nms_limit,
                         int proposal_count, float nms_threshold, const float* bbox_std_dev,
                         float* rpn_rois, int32_t* topk_idx, int32_t* keep_idx,
                         int32_t* keep_count, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- a8: PyramidROIAlign  (mrcnn/model.py:428-534, log2_graph :413-423)
 * feature_maps: HOST array of 4 device pointers P2..P5, each [B,feat_h[l],feat_w[l],C] of
 * `dtype`; boxes [B,N,4] float32 normalised -> pooled [B,N,P,P,C] (same dtype); levels [B,N]
 * int32 optional.  image_area = IMAGE_SHAPE[0]*IMAGE_SHAPE[1] as float32. */
int mrcnn_roi_levels(const float* boxes, int num_boxes, float image_area, int32_t* levels,
                     void* stream);
int mrcnn_pyramid_roi_align(const void* const* feature_maps, const int* feat_h, const int* feat_w,
                            int channels, int dtype, const float* boxes, int batch, int num_boxes,
                            int pool_size, float image_area, void* pooled, int32_t* levels,
                            void* stream);

/* ---- a10: DetectionLayer  (mrcnn/model.py:868-909, refine_detections_graph :770-865,
 *      norm_boxes_graph :3003-3017)
 * rois [B,N,4], mrcnn_class [B,N,NC], mrcnn_bbox [B,N,NC,4], image_metas [B,meta_size] float32
 * -> detections [B,D,6] = (y1,x1,y2,x2,class_id,score), zero padded.  min_confidence == 0
 * skips the confidence filter (reference truthiness test).  bbox_std_dev: HOST, 4 floats. */
int mrcnn_detection_layer(const float* rois, const float* mrcnn_class, const float* mrcnn_bbox,
                          const float* image_metas, int meta_size, int batch, int num_rois,
                          int num_classes, int max_instances, float min_confidence,
                          float nms_threshold, const float* bbox_std_dev, float* detections,
                          void* stream);

/* ---- a12: unmold_detections  (mrcnn/model.py:2558-2621; utils.norm_boxes / denorm_boxes /
 *      unmold_mask mrcnn/utils.py:923-954, 629-645)
 * detections [B,D,6], mrcnn_mask [B,D,MH,MW,NC] float32; orig_hw (HOST int[2]) = original image
 * size shared by the batch; image_hw (HOST int[2]) molded size; windows [B,4] int32 DEVICE
 * (pixel window of each image in the molded frame).
 * Outputs: rois [B,D,4] int32, class_ids [B,D] int32, scores [B,D] float32 (first counts[b] rows
 * valid, zero-area rows already removed), counts [B] int32, masks [B,H0,W0,D] uint8 (0/1), the
 * reference's [H,W,N] layout with N padded to D. */
int mrcnn_unmold_detections(const float* detections, const float* mrcnn_mask, int batch,
                            int max_instances, int mask_h, int mask_w, int num_classes,
                            const int* orig_hw, const int* image_hw, const int32_t* windows,
                            int32_t* rois, int32_t* class_ids, float* scores, int32_t* counts,
                            uint8_t* masks, void* workspace, size_t workspace_bytes, void* stream);
size_t mrcnn_unmold_workspace_bytes(int batch, int max_instances);
/* Same computation, masks as PIXEL-MAJOR BITS: mask_bits [B, H0*W0, DW] uint32, bit k of word w of a pixel = that
 * pixel of detection 32*w+k of the image (compacted order), DW = mrcnn_mask_bits_words(D) in {1,2,4,8}.  This is what the
 * engine ships over PCIe for host results (16 bytes per pixel for D = 100 instead of 100). */
int mrcnn_mask_bits_words(int max_instances);
int mrcnn_unmold_detections_bits(const float* detections, const float* mrcnn_mask, int batch,
                                 int max_instances, int mask_h, int mask_w, int num_classes,
                                 const int* orig_hw, const int* image_hw, const int32_t* windows,
                                 int32_t* rois, int32_t* class_ids, float* scores, int32_t* counts,
                                 uint32_t* mask_bits, void* workspace, size_t workspace_bytes, void* stream);
/* DEVICE counterpart of mrcnn_host_expand_mask_bits: mask_bits [n_images, pixels_per_image, DW] (DEVICE), counts
 * [n_images] (DEVICE) -> dense [n_images, pixels_per_image * max_instances] uint8 (DEVICE): the first pixels_per_image *
 * counts[i] bytes of slot i hold the C-order [pixels, counts[i]] 0/1 array (= masks [H0, W0, N] bool, mrcnn/model.py:
 * 2613-2619).  pixels_per_image * max_instances must be a multiple of 4. */
int mrcnn_mask_bits_expand_device(const uint32_t* mask_bits, const int32_t* counts, int n_images, int64_t pixels_per_image,
                                  int max_instances, uint8_t* dense, void* stream);
/* HOST-ONLY: expands pixel-major mask bits (HOST memory, as copied back from the device) into the reference's
 * result contract (mrcnn/model.py:2613-2619): for image b, dst[b] receives a dense C-order [pixels_per_image, counts[b]]
 * uint8 array (0/1) = masks [H0, W0, N] bool.  dst: HOST array of n_images HOST pointers (may be NULL where
 * counts[b] == 0).  n_threads <= 0: mrcnn_host_threads() workers (min(16, CPUs of the affinity mask),
 * env MRCNN_B200_HOST_THREADS overrides).  AVX-512BW when the CPU has it (env MRCNN_B200_HOST_SIMD=0 disables). */
int mrcnn_host_threads(void);
int mrcnn_host_expand_mask_bits(const uint32_t* bits, int n_images, int64_t pixels_per_image, int words_per_pixel,
                                const int32_t* counts, uint8_t* const* dst, int n_threads);

/* ---- a4-a6, a9, a11: the dense contractions — one bf16 tcgen05/TMEM implicit-GEMM family
 * (KL.Conv2D / TimeDistributed(Conv2D|Dense) / Conv2DTranspose call sites:
 *  mrcnn/model.py:99-210 backbone, :2003-2026 FPN, :916-957 RPN, :986-1039 class head,
 *  :1042-1091 mask head).
 * y[n,oh,ow,:] = act( scale * sum_{r,s,c} x[n, oh*stride+r-pad, ow*stride+s-pad, c] * w[:,r,s,c]
 *                     + shift + residual )
 * x: [N,H,W,Cin] bf16, Cin % 64 == 0; w: [Cout_pad, KH*KW*Cin] bf16 (K-major, Cout padded to the
 * N tile); scale/shift [Cout] float32 (folded BatchNorm + bias); residual optional bf16
 * [N,OH,OW,Cout] (or [N,OH/2,OW/2,Cout] with residual_upsample2 = nearest 2x); out bf16 or f32.
 * KHxKW in {1x1 (stride 1|2, pad 0), 3x3 (stride 1, pad 1)}; out_mode 1 = 2x2-stride-2 transposed
 * convolution scatter (Cout_total = 4*Cout taps-major). */
typedef struct mrcnn_conv_desc {
  int n, h, w, cin;          /* input tensor */
  int kh, kw, stride, pad;   /* filter geometry */
  int cout;                  /* real output channels (per tap for out_mode 1) */
  int relu;
  int residual_upsample2;
  int out_dtype;             /* MRCNN_DTYPE_* */
  int out_mode;              /* 0 = NHWC, 1 = deconv 2x2 s2 scatter */
  int out_ld;                /* output row pitch in elements (0 = cout) */
} mrcnn_conv_desc;
int mrcnn_conv2d_bf16(const mrcnn_conv_desc* desc, const void* x, const void* w, const float* scale,
                      const float* shift, const void* residual, void* out, void* stream);
/* Weight gradient of mrcnn_conv2d_bf16 for the training graph: dw[co, r, s, ci] += sum over pixels of
 * dy[n, oh, ow, co] * x[n, oh + r - pad, ow + s - pad, ci]  (float32 accumulation into dw, which is NOT cleared: it has
 * the [Cout, KH, KW, Cin] layout of the parameter and its gradient buffer).  x [N,H,W,Cin] bf16, dy [N,H,W,Cout] bf16;
 * 1x1 stride 1 or 3x3 stride 1 pad 1; Cin % 64 == 0, Cout % 8 == 0.  tcgen05 GEMM over MN-major operands, K (pixels)
 * split across CTAs.  Only n, h, w, cin, cout, kh, kw, stride, pad of the descriptor are read.  Optional (both or none):
 * w = the layer's bf16 weights [Cout, KH, KW, Cin] and wdot [Cout] float32, wdot[co] += <w[co], this call's dw[co]> — the
 * gradient of a per-channel scale applied behind the convolution (folded BatchNorm) times that scale. */
int mrcnn_conv2d_wgrad_bf16(const mrcnn_conv_desc* desc, const void* x, const void* dy, float* dw, const void* w,
                            float* wdot, void* stream);
/* Data gradient of mrcnn_conv2d_bf16 (train mode): dx[n,h,w,ci] = sum_{r,s,co} dy[n, h-r+pad, w-s+pad, co] * w[co,r,s,ci], the same
 * implicit GEMM with the B operand read MN-major from the FORWARD layer's weights w [Cout,KH,KW,Cin] (no transposed copy).
 * fwd describes the forward layer (n, h, w = the OUTPUT extent, which equals the input extent for these stride-1 layers);
 * dy [N,H,W,Cout] bf16 (already multiplied by any per-channel scale), dx [N,H,W,Cin] bf16; ones / zeros: device float32
 * vectors of at least Cin elements (epilogue scale / shift).  1x1 stride 1 or 3x3 stride 1 pad 1, Cin and Cout % 64 == 0. */
int mrcnn_conv2d_dgrad_bf16(const mrcnn_conv_desc* fwd, const void* dy, const void* w, const float* ones, const float* zeros,
                            void* dx, void* stream);
/* reference implementation on CUDA cores (fp32 accumulate) used only by the tests to check the
 * tcgen05 path on the device at full size */
int mrcnn_conv2d_bf16_simt(const mrcnn_conv_desc* desc, const void* x, const void* w, const float* scale,
                           const float* shift, const void* residual, void* out, void* stream);

/* ---- engine: keras_model.predict([molded_images, image_metas, anchors]) of mode='inference'
 *      (graph built at mrcnn/model.py:1935-2054, 2133-2159; outputs :2156-2158)
 *      + MaskRCNN.load_weights by layer name (mrcnn/model.py:2197-2239). */
typedef struct mrcnn_engine mrcnn_engine;

typedef struct mrcnn_engine_config {
  int batch_size;              /* config.BATCH_SIZE */
  int image_size;              /* config.IMAGE_SHAPE[0] == [1], multiple of 64 */
  int num_classes;             /* config.NUM_CLASSES */
  int pre_nms_limit;           /* config.PRE_NMS_LIMIT */
  int post_nms_rois;           /* config.POST_NMS_ROIS_INFERENCE */
  int detection_max_instances; /* config.DETECTION_MAX_INSTANCES */
  int pool_size;               /* config.POOL_SIZE */
  int mask_pool_size;          /* config.MASK_POOL_SIZE */
  int fc_layers_size;          /* config.FPN_CLASSIF_FC_LAYERS_SIZE */
  int top_down_pyramid_size;   /* config.TOP_DOWN_PYRAMID_SIZE */
  int anchors_per_location;    /* len(config.RPN_ANCHOR_RATIOS) */
  float rpn_nms_threshold;     /* config.RPN_NMS_THRESHOLD */
  float detection_min_confidence;
  float detection_nms_threshold;
  float rpn_bbox_std_dev[4];
  float bbox_std_dev[4];
  int backbone_strides[5];
} mrcnn_engine_config;

int mrcnn_engine_create(const mrcnn_engine_config* cfg, int device, mrcnn_engine** out);
void mrcnn_engine_destroy(mrcnn_engine* e);
/* number of weighted layers the graph expects, and their names / kinds / shapes */
int mrcnn_engine_num_layers(const mrcnn_engine* e);
int mrcnn_engine_layer_info(const mrcnn_engine* e, int index, const char** name, int* kind /*0 conv,1 bn,2 dense,3 deconv*/,
                            int* num_weights, int* shape4 /* kernel shape, padded with 0 */);
/* by-name weight load; weight_index follows Keras layer.weights order (kernel,bias | gamma,beta,
 * mean,var); data: HOST float32 in the Keras layout; count = number of floats (checked). */
int mrcnn_engine_set_weight(mrcnn_engine* e, const char* layer_name, int weight_index,
                            const float* data, size_t count);
/* fold BN, convert to bf16 GEMM layout, upload.  Fails listing the first missing layer unless
 * allow_missing (missing layers keep zeros). */
int mrcnn_engine_finalize(mrcnn_engine* e, int allow_missing);
/* anchors [A,4] float32 normalised (HOST), as MaskRCNN.get_anchors returns them */
int mrcnn_engine_set_anchors(mrcnn_engine* e, const float* anchors, int num_anchors);
/* Runs the graph.  molded [B,S,S,3] float32 and image_metas [B,12+NC] float32: HOST pointers when
 * inputs_on_host != 0 (copied H2D on the engine stream), else device pointers.  Blocks until the
 * outputs are resident (device) unless async != 0. */
int mrcnn_engine_predict(mrcnn_engine* e, const float* molded, const float* image_metas,
                         int inputs_on_host, int async);
/* device pointers of the graph outputs / taps, valid until the next predict:
 * "detections" [B,D,6] f32, "mrcnn_class" [B,R,NC] f32, "mrcnn_bbox" [B,R,NC,4] f32, "mrcnn_mask"
 * [B,D,28,28,NC] f32, "rpn_rois" [B,R,4] f32, "rpn_class" [B,A,2] f32, "rpn_bbox" [B,A,4] f32,
 * "P2".."P6" bf16 NHWC, "C2".."C5" bf16, "pooled" / "pooled_mask" bf16, "topk_idx", "keep_idx",
 * "keep_count", "roi_levels" int32.  Returns NOTFOUND for unknown names. */
int mrcnn_engine_tensor(const mrcnn_engine* e, const char* name, void** device_ptr, size_t* bytes);
/* copies a named tensor to a HOST buffer as float32 (bf16 tensors are widened) or int32 */
int mrcnn_engine_read(const mrcnn_engine* e, const char* name, void* host_dst, size_t dst_bytes);
/* run individual stages on engine-owned tensors from caller-provided DEVICE inputs ("reference-fed"
 * microbenchmarks, BASELINE config #3): stage in {"proposal","roialign","class_head",
 * "detection","roialign_mask","mask_head"}; consumes the current contents of the engine's input
 * tensors for that stage (writable through mrcnn_engine_write). */
int mrcnn_engine_run_stage(mrcnn_engine* e, const char* stage);
int mrcnn_engine_write(mrcnn_engine* e, const char* name, const void* host_src, size_t src_bytes);
/* whole detect(): predict + unmold, results to HOST buffers (pinned recommended).
 * molded [B,S,S,3] float32 (MaskRCNN.mold_inputs output): HOST pointer when molded_on_host != 0,
 * else DEVICE; metas / windows (int32 [B,4]) / orig_hw: HOST.  Outputs as for
 * mrcnn_unmold_detections_bits, written to HOST buffers; mask_bits [B,H0*W0,DW] uint32 (expand with
 * mrcnn_host_expand_mask_bits).  Blocking. */
int mrcnn_engine_detect_molded(mrcnn_engine* e, const float* molded, int molded_on_host,
                               const float* metas_host, const int* orig_hw,
                               const int32_t* windows_host, int32_t* rois_host,
                               int32_t* class_ids_host, float* scores_host, int32_t* counts_host,
                               uint32_t* mask_bits_host);
/* Hybrid delivery of the host masks (optional, applies to the NEXT detect call that returns host results, then resets):
 * the masks of images [0, n_images) are ALSO expanded on the device (mrcnn_mask_bits_expand_device) and copied by the DMA
 * engine into dense_host (PINNED; slot i at dense_host + i * H0*W0*DETECTION_MAX_INSTANCES, dense [H0*W0, counts[i]] bytes),
 * so the caller's cores only expand the mask bits of the remaining images.  The bits of every image are still delivered.
 * mrcnn_engine_dense_copy_ms: after mrcnn_engine_wait_slot / a blocking call, the number of images delivered this way for
 * that result slot and the device time of expansion + copy (for balancing the share against the host's expansion rate). */
int mrcnn_engine_set_dense_output(mrcnn_engine* e, uint8_t* dense_host, int n_images);
/* like mrcnn_engine_wait_slot, but returns as soon as everything except the dense share is on the host (boxes, ids, scores,
 * counts, mask bits), so that the caller can expand the other images while the DMA copy of the dense share is still in
 * flight; mrcnn_engine_wait_slot then waits for the rest. */
int mrcnn_engine_wait_slot_packed(mrcnn_engine* e, int slot);
int mrcnn_engine_dense_copy_ms(mrcnn_engine* e, int slot, int* n_images, float* ms);
/* The whole hot path from FITS-like maps in ONE call: maps [B,map_h,map_w] float32 (HOST when
 * maps_on_host != 0, else DEVICE; NaN allowed) -> zscale/uint8 RGB (read_fits, mrcnn/utils.py:
 * 1090-1208) -> resize to (out_h,out_w), pad at (top,left), minus mean (mold_inputs, mrcnn/model.py:
 * 2519-2556) -> graph -> unmold against the original (map_h,map_w) frame.  metas / windows: HOST.
 * mask_format: 1 = full-frame masks as pixel-major bits (device tensor "unmold_mask_bits", copied to mask_bits_host
 * [B,H0*W0,DW] uint32 when that pointer is given); 0 = [B,H0,W0,D] uint8 kept on the device only (tensor "unmold_masks",
 * consumed by the mrcnn.analyze bit-plane kernels; mask_bits_host must be NULL).
 * Host result pointers may be NULL (all of them = results stay on the device, readable through
 * mrcnn_engine_tensor "unmold_rois"/"unmold_class_ids"/"unmold_scores"/"unmold_counts"/"unmold_masks"|"unmold_mask_bits";
 * the result slot of a call is mrcnn_engine_next_slot() read before it, slot 1 names carry the suffix "#1").
 * Blocking unless async. */
int mrcnn_engine_detect_maps(mrcnn_engine* e, const float* maps, int maps_on_host, int map_h, int map_w,
                             const float* contrasts3, const float* mean_pixel3, int out_h, int out_w,
                             int top, int left, const float* metas_host, const int32_t* windows_host,
                             int32_t* rois_host, int32_t* class_ids_host, float* scores_host,
                             int32_t* counts_host, uint32_t* mask_bits_host, int mask_format, int async);
/* async != 0 (with host result pointers): returns once the work is queued; the device->host copies
 * run on a second stream from double-buffered result slots, so the next call's compute overlaps
 * them.  At most two calls may be in flight; mrcnn_engine_wait() blocks until every queued copy
 * has landed in its host buffers (which must stay valid and pinned until then). */
int mrcnn_engine_wait(mrcnn_engine* e);
/* slot (0/1) the next async call will use, and a wait for the copies of one slot only */
int mrcnn_engine_next_slot(const mrcnn_engine* e);
int mrcnn_engine_wait_slot(mrcnn_engine* e, int slot);
void* mrcnn_engine_stream(const mrcnn_engine* e);
/* per-stage device time of the last predict in milliseconds (CUDA events); names via index */
int mrcnn_engine_stage_times(const mrcnn_engine* e, int max_stages, const char** names, float* ms);
/* Per-kernel-family device time of the last predict (CUDA events recorded around every launch
 * on the engine stream while profiling is enabled): returns the number of families; names[k] in
 * {"conv_gemm","roialign","proposal","detection","stem_im2col","maxpool",...}. */
/* enable: 0 off; 1 events around every launch (per-launch table via mrcnn_engine_step_info, times of the last
 * predict); 2 events only where the kernel family changes between consecutive launches (~35 instead of ~150
 * per predict), kept for the last 8 predicts so a timed loop needs no sync: kernel_times then returns the
 * per-predict average over those. */
int mrcnn_engine_set_profiling(mrcnn_engine* e, int enable);
int mrcnn_engine_kernel_times(mrcnn_engine* e, int max_kinds, const char** names, float* ms, int* launches);
/* one launch of the plan: label (output tensor), kernel family, device ms of the last profiled
 * predict (-1 if none) and GEMM FLOPs; returns NOTFOUND past the last launch */
int mrcnn_engine_step_info(mrcnn_engine* e, int index, const char** label, const char** kind, float* ms, double* flops);
/* FLOPs (2*M*N*K over all GEMM launches) of one predict at the configured batch */
double mrcnn_engine_flops(const mrcnn_engine* e);

/* ---- (f1) Analyzer source-mask post-processing on device-resident masks
 *      (mrcnn/analyze.py: Analyzer.extract_det_masks :1162-1423, make_json_results :1866-1942,
 *       merge_masks :2142-2146, extract_mask_connected_components :2148-2151, are_mask_connected
 *       :2154-2173; utils.extract_bboxes mrcnn/utils.py:33-59).
 * Masks are handled as bit-planes: plane[m] is [height][ceil(width/32)] uint32, bit k of word w is
 * pixel x = 32*w + k, bits past `width` are zero.  mrcnn_plane_words() = words per plane.  All
 * pointers are DEVICE pointers unless stated; calls are asynchronous on `stream`.  The graph logic
 * between the calls (score filter, merge graph, cliques) is host logic in mrcnn/analyze.py. */
size_t mrcnn_plane_words(int height, int width);
/* masks [n_images,H,W,depth] uint8 (0 / non-zero; the [H,W,N] result layout of detect with N padded
 * to depth); plane_of [n_images*depth] int32: destination plane of detection (b,d) or -1 = skip. */
int mrcnn_masks_pack(const uint8_t* masks, int n_images, int height, int width, int depth,
                     const int32_t* plane_of, uint32_t* planes, void* stream);
/* area [n] = pixel count; bbox [n,4] = (y1,x1,y2,x2) of utils.extract_bboxes (y2/x2 exclusive; zeros
 * for an empty mask). */
int mrcnn_planes_area_bbox(const uint32_t* planes, int n_planes, int height, int width, int32_t* area,
                           int32_t* bbox, void* stream);
/* pairs [n_pairs,2] plane indices -> inter[p] = |a & b| (numerator of sklearn jaccard_score, the
 * denominator is area[a]+area[b]-inter), touch[p] = 1 iff are_mask_connected(a,b) would be True
 * (a pixel of a coincides with or is 4-adjacent to a pixel of b).  bbox: optional [n_planes,4] from
 * mrcnn_planes_area_bbox (NULL = none): pairs whose boxes are more than one pixel apart are answered without
 * reading the planes, the others scan only the rows of a's box. */
int mrcnn_planes_pair_stats(const uint32_t* planes, int height, int width, const int32_t* pairs, int n_pairs,
                            const int32_t* bbox, int32_t* inter, int32_t* touch, void* stream);
/* out[g] = OR of planes[members[offsets[g] .. offsets[g+1])]  (merge_masks folded over a group) */
int mrcnn_planes_union(const uint32_t* planes, int height, int width, const int32_t* members,
                       const int32_t* offsets, int n_groups, uint32_t* out, void* stream);
/* skimage.measure.label(mask, background=0, connectivity=1): labels [n,H,W] int32 (0 = background,
 * 1.. numbered in raster order of each component's first pixel), counts [n] = ncomponents. */
size_t mrcnn_planes_label_workspace_bytes(int n_planes, int height, int width);
int mrcnn_planes_label(const uint32_t* planes, int n_planes, int height, int width, int32_t* labels,
                       int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
/* planes_out[k] = (labels[src[k]] == comp[k])  (np.where(component_labels == i+1, [1], [0])) */
int mrcnn_labels_select(const int32_t* labels, int height, int width, const int32_t* src, const int32_t* comp,
                        int n_out, uint32_t* planes_out, void* stream);
/* np.argwhere(mask == 1) of every plane, row-major: pixels [total,2] int32 = (y + y_origin,
 * x + x_origin); plane m writes rows offsets[m] .. offsets[m]+area[m] (offsets: int64 [n]). */
int mrcnn_planes_pixels(const uint32_t* planes, int n_planes, int height, int width, const int64_t* offsets,
                        int y_origin, int x_origin, int32_t* pixels, void* stream);
/* out [n,H,W] uint8 0/1 */
int mrcnn_planes_unpack(const uint32_t* planes, int n_planes, int height, int width, uint8_t* out, void* stream);

/* ---- (f2) tile driver: merging of sources across tile edges (mrcnn/sfinder.py: SFinder.merge_edge_sources
 *      :711-935, pixel test :787-808).  pixels [total,2] int32 (y,x) = the concatenated pixel lists of all edge sources
 *      in one global frame, offsets int64 [n_lists+1]; pairs [n_pairs,2] list indices (candidates that already passed
 *      the neighbour-tile and bounding-box tests) -> adjacent[p] = 1 iff some pixel of list a and some pixel of list b
 *      satisfy |dx| <= 1 and |dy| <= 1 (8-connected, the reference's double loop). */
int mrcnn_pixel_lists_adjacent(const int32_t* pixels, const int64_t* offsets, const int32_t* pairs, int n_pairs,
                               int32_t* adjacent, void* stream);

/* HOST-ONLY (no device work): the merge graph of Analyzer.extract_det_masks (mrcnn/analyze.py:1258-1320, mrcnn/graph.py)
 * for a batch of frames.  det_count [n_frames] masks per frame (global indices are frame-major); pairs [n_pairs,2]
 * global indices in the reference's pair-loop order, mergeable [n_pairs] (0/1).  Output, all HOST: members [sum
 * det_count] and offsets [sum det_count + 1] = the connected components in the reference's order (by smallest vertex;
 * members in recursive depth-first pre-order over insertion-ordered adjacency lists), frame_components [n_frames] =
 * components per frame, *n_components. */
int mrcnn_host_merge_components(int n_frames, const int32_t* det_count, const int32_t* pairs, const uint8_t* mergeable,
                                int64_t n_pairs, int32_t* members, int32_t* offsets, int32_t* frame_components,
                                int32_t* n_components);

/* HOST-ONLY: all (i < j) pairs inside every frame, frame-major, in the reference's double-loop order
 * (mrcnn/analyze.py:1263-1266 and :1335-1338); pairs [sum n_f(n_f-1)/2, 2] global indices. */
int mrcnn_host_all_pairs(int n_frames, const int32_t* counts, int32_t* pairs);

/* HOST-ONLY: the per-pair decisions of extract_det_masks for a batch, pairs implicit in mrcnn_host_all_pairs order, from
 * the device's pair statistics (inter, touch) and plane areas.  stage 0 = merge graph edges (mrcnn/analyze.py:1268-1290:
 * connected, same class, IOU >= iou_thr; cls_or_spurious = class ids).  stage 1 = selection graph edges (:1340-1360:
 * connected, and with use_iou (split_source_sidelobe) not a spurious/non-spurious pair below iou_thr; cls_or_spurious =
 * 1 for 'spurious'), plus loses[v] = 1 when a linked mask scores strictly higher (caller zero-fills) and tie_frame[f] = 1
 * when two linked masks of frame f score equal.  IOU as sklearn's jaccard_score: float64, 0 for an empty union. */
int mrcnn_host_pair_flags(int stage, int n_frames, const int32_t* counts, const int32_t* cls_or_spurious, const float* score,
                          const int32_t* area, const int32_t* inter, const int32_t* touch, int use_iou, double iou_thr,
                          uint8_t* flags, uint8_t* loses, uint8_t* tie_frame);

/* HOST-ONLY: the `vertexes` of catalogue objects (mrcnn/analyze.py:1908-1927, mrcnn/sfinder.py:885-910):
 * skimage.measure.find_contours(zero-padded mask, 0.5) [scikit-image 0.15 algorithm, third-party, restated; parity
 * unpinned] computed from each object's pixel list.  pixels_yx [total,2] int32 (y, x) in image coordinates,
 * pixel_offsets int64 [n_objects+1].  The result stays in thread-local storage: n_vertices / n_contours give its size,
 * mrcnn_host_contours_fetch copies it out: vertices_xy [n_vertices,2] float64 (x, y) in the same coordinates
 * (= pixel +- 0.5), contour_offsets int64 [n_contours+1] (vertex ranges, contours in skimage's order),
 * object_offsets int64 [n_objects+1] (contour ranges). */
int mrcnn_host_contours(const int32_t* pixels_yx, const int64_t* pixel_offsets, int n_objects, int64_t* n_vertices,
                        int64_t* n_contours);
int mrcnn_host_contours_fetch(double* vertices_xy, int64_t* contour_offsets, int64_t* object_offsets);

/* ---- train mode (BASELINE.json configs[4]; reference MaskRCNN(mode='training'), mrcnn/model.py:2068-2132) -----------
 * The dense layers of the training graph are driven from mrcnn/training.py; these are its graph layers that are not
 * contractions, plus the optimiser. */

/* DetectionTargetLayer (mrcnn/model.py:570-763): per image, trim zero proposals / GT rows, IoU proposals x GT,
 * positives (IoU >= 0.5) and negatives (< 0.5, not in a crowd box), shuffle + sub-sample to
 * int(train_rois * roi_positive_ratio) positives and the matching negatives, box refinement targets / bbox_std_dev, and the
 * mask_h x mask_w crop_and_resize (rounded) of the assigned GT mask.  gt_masks [B, mask_src_h, mask_src_w, max_gt] 0/1
 * bytes (full-size masks, or mini masks with use_mini_mask = 1).  tf.random.shuffle is replaced by ascending
 * mrcnn_shuffle_key(seed, image, stream 0 positives / 1 negatives, trimmed proposal index); seed_device (optional, device
 * memory) is added to seed when the kernel runs, so a replayed CUDA graph can advance it.  Outputs are zero padded to
 * train_rois rows: rois [B,T,4], target_class_ids [B,T], target_bbox [B,T,4], target_mask [B,T,mask_h,mask_w];
 * counts (optional) [B,2] = positives, negatives. */
uint32_t mrcnn_shuffle_key(unsigned long long seed, uint32_t image, uint32_t stream, uint32_t index);
int mrcnn_detection_targets(const float* proposals, const int32_t* gt_class_ids, const float* gt_boxes,
                            const uint8_t* gt_masks, int batch, int num_proposals, int max_gt, int mask_src_h,
                            int mask_src_w, int use_mini_mask, int train_rois, float roi_positive_ratio,
                            const float* bbox_std_dev, int mask_h, int mask_w, unsigned long long seed,
                            const unsigned long long* seed_device, float* rois,
                            int32_t* target_class_ids, float* target_bbox, float* target_mask, int32_t* counts,
                            void* stream);

/* Gradient of mrcnn_pyramid_roi_align (mrcnn/model.py:428-534) with respect to the pyramid: dpooled [B,N,P,P,C] bf16 is
 * scattered with the forward's bilinear weights into dfeature_maps[l] [B,H_l,W_l,C] float32 (accumulated: clear them
 * first); levels [B,N] are the forward's ROI levels. */
int mrcnn_pyramid_roi_align_backward(float* const* dfeature_maps, const int* feat_h, const int* feat_w, int channels,
                                     const float* boxes, const int32_t* levels, int batch, int num_boxes, int pool_size,
                                     const void* dpooled_bf16, void* stream);

/* First step of the backward pass of a fused conv + affine (+ residual) + ReLU layer, one pass over the gradient:
 * dz = dout * (out > 0) (out = the layer's forward output; NULL: no ReLU), dzs = dz * scale[c] (scale NULL: dzs = dz),
 * colsum[c] += sum over rows of dz.  [rows, channels] bf16 row-major, channels % 8 == 0, <= 2048; dz / dzs / colsum optional. */
int mrcnn_conv_backward_prep(const void* dout, const void* out, const float* scale, void* dz, void* dzs, float* colsum,
                             long long rows, int channels, void* stream);

/* One optimiser step over flat float32 buffers of n elements (mrcnn/model.py:2259-2297: keras SGD(lr, momentum,
 * clipnorm=GRADIENT_CLIP_NORM) on loss + sum_w l2(WEIGHT_DECAY)(w) / size(w)):
 *   g = grad * grad_scale + segment_reg_coef[s] * w      (s = segment of the element; coef = 2*WEIGHT_DECAY/size(w), 0 for BN)
 *   g *= clipnorm / max(||g||_2 over ALL n elements, clipnorm)      (clipnorm <= 0: no clipping)
 *   velocity = momentum * velocity - learning_rate * g ;  weights += velocity ;  weights_bf16 (optional) = bf16(weights)
 * segment_start int64 [num_segments + 1] (device), sumsq_scratch: one device double. */
int mrcnn_sgd_step(float* grad, float* weights, float* velocity, void* weights_bf16, long long n,
                   const long long* segment_start, const float* segment_reg_coef, int num_segments, float grad_scale,
                   float clipnorm, float learning_rate, float momentum, double* sumsq_scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRCNN_B200_H_ */
