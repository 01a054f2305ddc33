// Internal C++ interface of the bf16 implicit-GEMM convolution family (conv_gemm.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "mrcnn_b200.h"

struct ConvGemmParams {
  // M tiling: an M tile is a TMA box (64ch, tw, th, nb) of the (possibly strided) input view
  int tiles_w, tiles_h, tiles_nb, n_tiles;
  int tw, th, nb;
  int OW, OH, N;      // output spatial size and image count
  int flat;           // 1x1 stride-1 (and im2col mode): M tiles run over the flattened pixel index
  int im2col;         // 3x3: A tiles are fetched with TMA im2col-mode loads (128 consecutive output pixels)
  long long M;        // N*OH*OW
  int kh, kw, pad;    // filter taps (stride is folded into the tensor-map view)
  int cin_blocks;     // Cin / 64
  int cout;           // real output channels (per tap in deconv mode)
  int cout_total;     // GEMM N extent (4*cout in deconv mode)
  int relu, res_up2, out_f32, out_mode, out_ld;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
  // out_mode 2 (engine-internal): 2x2-s2 transposed conv + ReLU fused with the following 1x1 conv
  // (cout -> nc2) + sigmoid; `out` is then float32 [N, 2*OH, 2*OW, nc2] (mrcnn_mask)
  const __nv_bfloat16* w2;   // [nc2][cout] bf16
  const float* b2;           // [nc2]
  int nc2;
  int unit_scale;            // out_mode 2: every per-channel scale is exactly 1 (no BN on the transposed conv)
  // optional (flat / im2col layers without residual): tile_skip[m_tile] != 0 -> this M tile is not computed and its output
  // rows are left untouched (mask head: tiles that hold only zero-padded detections).  Device memory written by an earlier
  // kernel of the stream; every role of the kernel reads the same flags, so ring / accumulator phases stay in step.
  const unsigned char* tile_skip;
};

struct ConvPlan {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  CUtensorMap tmap_out;        // epi_tma: [M, cout] bf16 output, box 32 ch x 128 rows, 64B swizzle
  CUtensorMap tmap_res;        // epi_tma: same geometry over the residual tensor
  int epi_tma = 0;             // epilogue through shared memory with TMA residual loads / output stores
  int occ2 = 0;                // launch attribute (may be set after conv_plan_create): two persistent CTAs per SM with a shorter
                               // ring (block_n <= 128, forward / K-major B only); same results
  int b_mn = 0;                // set BEFORE conv_plan_create: data-gradient mode, B read MN-major from the forward weights
  ConvGemmParams p;
  int block_n;
  dim3 grid;
  double flops;
  float* w2_table = nullptr;   // fused mask logits: [256][8] float expansion of w2 (device; freed by the owner)
};

// Builds the tensor maps + launch geometry for one layer.  x/w/out are device pointers that must
// stay valid for the life of the plan.  block_n = 0 picks a tile width from cout.
int conv_plan_create(const mrcnn_conv_desc* d, const void* x, const void* w, const float* scale,
                     const float* shift, const void* residual, void* out, int block_n, ConvPlan* plan);
// epi_tma: -1 = policy, 0 = direct stores, 1 = TMA epilogue when the layer is eligible (results are bit-identical)
int conv_plan_create_ex(const mrcnn_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                        const void* residual, void* out, int block_n, int epi_tma, ConvPlan* plan);
bool conv_plan_epi_tma_eligible(const mrcnn_conv_desc* d);
int conv_plan_launch(const ConvPlan* plan, cudaStream_t stream);
// turns a 256-channel deconv plan (block_n 256) into deconv + ReLU + 1x1 conv (nc2) + sigmoid -> float32 out
int conv_plan_fuse_mask_logits(ConvPlan* plan, const void* w2, const float* b2, int nc2, void* out, int unit_scale);

// flags [m_tiles] (see ConvGemmParams::tile_skip); returns an error for plans whose M tiles are not 128 consecutive rows
int conv_plan_set_tile_skip(ConvPlan* plan, const unsigned char* flags);

void mrcnn_count_launch(unsigned long long n);
