// Launchers of the memory-bound glue kernels (elementwise.cu).
#pragma once
#include "common.cuh"

void mrcnn_count_launch(unsigned long long n);

int launch_stem_im2col(const float* img, int B, int S, __nv_bfloat16* A, cudaStream_t st);
int launch_maxpool3x3s2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st);
int launch_subsample2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st);
int launch_rpn_post(const float* head, int ld, int B, int hw, int apl, int A, int level_off, float* rpn_class,
                    float* rpn_bbox, cudaStream_t st);
int launch_class_post(const float* head, int ld, int M, int NC, float* probs, float* bbox, cudaStream_t st);
int launch_mask_post(const float* logits, int ld, size_t M, int NC, float* out, cudaStream_t st);
// mask head work list: flags[t] = 1 for M tiles holding only zero-padded detections; zeroing of their mrcnn_mask rows
int launch_mask_tile_flags(const float* det, int n_rois, int rows_per_roi, int tile_rows, int n_tiles, unsigned char* flags,
                           cudaStream_t st);
int launch_mask_zero_padded(const float* det, int n_rois, size_t floats_per_roi, float* out, cudaStream_t st);
