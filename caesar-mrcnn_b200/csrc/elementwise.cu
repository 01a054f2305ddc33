// Memory-bound glue kernels of the dense graph (all NHWC, 16-byte vector accesses):
//   stem_im2col   : ZeroPadding2D(3) + 7x7 stride-2 patch gather for conv1 (mrcnn/model.py:183-184)
//   maxpool3x3s2  : MaxPooling2D((3,3), strides 2, padding="same") with TF SAME asymmetry (:187)
//   subsample2    : MaxPooling2D(pool 1, strides 2) = x[::2, ::2] for P6 (:2022)
//   rpn_post      : reshape [B,-1,2] + softmax, reshape [B,-1,4], concat over levels (:938-955, :2044-2053)
//   class_post    : softmax over classes + reshape of bbox deltas (:1028-1037)
//   mask_post     : sigmoid of the mask logits (:1088-1090)
#include "elementwise.cuh"

namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A[(b*OH+oh)*OW+ow][k], k = r*24 + s*3 + c (s*3+c < 21; the 3 tail entries of each 24-group and k >= 168 are
// zero; KPAD = 192).  A filter row of a patch is 21 contiguous floats of the NHWC image, so one CTA stages the 7
// zero-padded input rows of one output row in shared memory as bf16 (each pixel converted once) and every thread
// then emits aligned 16-byte vectors: 4 x LDS.32 + one coalesced STG.128, no per-element index arithmetic.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, int B, int S, int RW,
                                                          __nv_bfloat16* __restrict__ A) {
  pdl_prologue();
  extern __shared__ uint32_t stem_smem[];                       // [7][RW] bf16, RW even
  __nv_bfloat16* srow = reinterpret_cast<__nv_bfloat16*>(stem_smem);
  const int OH = S / 2, OW = S / 2;
  const int b = blockIdx.x / OH, oh = blockIdx.x - b * OH;
  const int n4 = S * 3 / 4;                                     // float4 chunks per image row
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int i = threadIdx.x; i < 7 * RW; i += blockDim.x) {      // padding (and rows outside the image)
    const int r = i / RW, q = i - r * RW;
    const int ih = oh * 2 + r - 3;
    if (q < 9 || q >= 9 + 3 * S || ih < 0 || ih >= S) srow[i] = zero;
  }
  for (int i = threadIdx.x; i < 7 * n4; i += blockDim.x) {
    const int r = i / n4, c4 = i - r * n4;
    const int ih = oh * 2 + r - 3;
    if (ih < 0 || ih >= S) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + ((size_t)b * S + ih) * S * 3) + c4);
    __nv_bfloat16* d = srow + r * RW + 9 + c4 * 4;              // odd element offset: 2 + 4 + 2 byte stores
    d[0] = __float2bfloat16_rn(v.x);
    *reinterpret_cast<uint32_t*>(d + 1) = pack2(v.y, v.z);
    d[3] = __float2bfloat16_rn(v.w);
  }
  __syncthreads();
  const uint32_t* sw = stem_smem;
  const int RW2 = RW / 2;
  __nv_bfloat16* arow = A + (size_t)blockIdx.x * OW * 192;
  for (int idx = threadIdx.x; idx < OW * 24; idx += blockDim.x) {
    const int ow = idx / 24, v = idx - ow * 24;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (v < 21) {
      const int r = v / 3, part = v - r * 3;
      const uint32_t* src = sw + r * RW2 + ow * 3 + part * 4;   // element ow*6 + part*8
      o.x = src[0];
      o.y = src[1];
      o.z = src[2];
      o.w = src[3];
      if (part == 2) {                                          // elements 21..23 of the group are padding
        o.z &= 0x0000ffffu;
        o.w = 0u;
      }
    }
    *reinterpret_cast<uint4*>(arow + (size_t)idx * 8) = o;
  }
}

__device__ __forceinline__ void max8(uint4& acc, const uint4 v) {
  __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&acc);
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], b[i]);
}

__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C,
                                    __nv_bfloat16* __restrict__ y) {
  pdl_prologue();
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const int pt_h = max((OH - 1) * 2 + 3 - H, 0) / 2, pt_w = max((OW - 1) * 2 + 3 - W, 0) / 2;
  const int cv = C / 8;
  const size_t total = (size_t)B * OH * OW * cv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv);
    size_t pix = idx / cv;
    const int ow = (int)(pix % OW);
    const int oh = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((size_t)OW * OH));
    uint4 acc;
    bool first = true;
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - pt_h + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 - pt_w + s;
        if (iw < 0 || iw >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + ih) * W + iw) * C + c8 * 8));
        if (first) { acc = v; first = false; } else max8(acc, v);
      }
    }
    *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) = acc;
  }
}

__global__ void subsample2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C,
                                  __nv_bfloat16* __restrict__ y) {
  pdl_prologue();
  const int OH = (H + 1) / 2, OW = (W + 1) / 2, cv = C / 8;
  const size_t total = (size_t)B * OH * OW * cv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv);
    size_t pix = idx / cv;
    const int ow = (int)(pix % OW);
    const int oh = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((size_t)OW * OH));
    *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) =
        __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + 2 * oh) * W + 2 * ow) * C + c8 * 8));
  }
}

// head [B*hw, ld] f32: cols [0, 2*apl) class logits (anchor a -> 2a,2a+1), cols [2*apl, 6*apl) deltas (4a+k)
__global__ void rpn_post_kernel(const float* __restrict__ head, int ld, int B, int hw, int apl, int A, int level_off,
                                float* __restrict__ rpn_class, float* __restrict__ rpn_bbox) {
  pdl_prologue();
  const int total = B * hw * apl;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int a = idx % apl;
    const int pix = (idx / apl) % hw;
    const int b = idx / (apl * hw);
    const float* row = head + ((size_t)b * hw + pix) * ld;
    const float l0 = row[2 * a], l1 = row[2 * a + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv = 1.0f / (e0 + e1);
    const size_t anchor = (size_t)b * A + level_off + (size_t)pix * apl + a;
    *reinterpret_cast<float2*>(rpn_class + anchor * 2) = make_float2(e0 * inv, e1 * inv);
    const float* d = row + 2 * apl + 4 * a;
    *reinterpret_cast<float4*>(rpn_bbox + anchor * 4) = make_float4(d[0], d[1], d[2], d[3]);
  }
}

// head [M, ld] f32: cols [0,NC) logits, cols [NC, 5*NC) bbox deltas (4*cls+k)
__global__ void class_post_kernel(const float* __restrict__ head, int ld, int M, int NC, float* __restrict__ probs,
                                  float* __restrict__ bbox) {
  pdl_prologue();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const float* row = head + (size_t)i * ld;
    float m = row[0];
    for (int k = 1; k < NC; ++k) m = fmaxf(m, row[k]);
    float sum = 0.f;
    for (int k = 0; k < NC; ++k) sum += expf(row[k] - m);
    const float inv = 1.0f / sum;
    for (int k = 0; k < NC; ++k) probs[(size_t)i * NC + k] = expf(row[k] - m) * inv;
    for (int k = 0; k < 4 * NC; ++k) bbox[(size_t)i * 4 * NC + k] = row[NC + k];
  }
}

__global__ void mask_post_kernel(const float* __restrict__ logits, int ld, size_t M, int NC, float* __restrict__ out) {
  pdl_prologue();
  const size_t total = M * NC;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / NC;
    const int k = (int)(idx - r * NC);
    const float v = logits[r * ld + k];
    out[idx] = 1.0f / (1.0f + expf(-v));
  }
}

// Mask-head work list (engine-internal).  The reference runs the mask branch on all DETECTION_MAX_INSTANCES rows of
// `detections`, zero padding included (mrcnn/model.py:2118-2132), and unmold_detections then drops every row from the first
// class id 0 on (:2575-2577).  flags[t] = 1 when M tile t (tile_rows consecutive rows of the [B*D * rows_per_roi] GEMMs of
// the mask head) holds rows of padded detections only: those tiles are never computed.
__global__ void mask_tile_flags_kernel(const float* __restrict__ det, int n_rois, int rows_per_roi, int tile_rows, int n_tiles,
                                       unsigned char* __restrict__ flags) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const long long r0 = (long long)t * tile_rows;
  int lo = (int)(r0 / rows_per_roi), hi = (int)((r0 + tile_rows - 1) / rows_per_roi);
  if (hi > n_rois - 1) hi = n_rois - 1;
  bool any = false;
  for (int r = lo; r <= hi; ++r) any = any || (det[(size_t)r * 6 + 4] != 0.f);
  flags[t] = any ? 0 : 1;
}

// rows of mrcnn_mask that belong to padded detections: written as zeros (their tiles may have been skipped)
__global__ void mask_zero_padded_kernel(const float* __restrict__ det, size_t floats_per_roi, float* __restrict__ out) {
  pdl_prologue();
  const int roi = blockIdx.x;
  if (det[(size_t)roi * 6 + 4] != 0.f) return;
  float4* dst = reinterpret_cast<float4*>(out + (size_t)roi * floats_per_roi);
  for (size_t i = threadIdx.x; i < floats_per_roi / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int launch_stem_im2col(const float* img, int B, int S, __nv_bfloat16* A, cudaStream_t st) {
  MRCNN_REQUIRE(S % 4 == 0, "stem_im2col: image size must be a multiple of 4");
  const int RW = ((S + 6) * 3 + 7) & ~7;
  const size_t smem = (size_t)7 * RW * 2;
  MRCNN_REQUIRE(smem <= 200 * 1024, "stem_im2col: image rows of %d pixels do not fit in shared memory", S);
  if (smem > 48 * 1024) {
    static size_t attr_bytes[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && attr_bytes[dev] < smem) {
      MRCNN_CHECK_CUDA(cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_bytes[dev] = smem;
    }
  }
  MRCNN_CHECK_CUDA(mrcnn_launch(stem_im2col_kernel, dim3(B * (S / 2)), dim3(256), smem, st, img, B, S, RW, A));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_maxpool3x3s2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st) {
  const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  MRCNN_CHECK_CUDA(mrcnn_launch(maxpool3x3s2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, x, B, H, W, C, y));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_subsample2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st) {
  const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  MRCNN_CHECK_CUDA(mrcnn_launch(subsample2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, x, B, H, W, C, y));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_rpn_post(const float* head, int ld, int B, int hw, int apl, int A, int level_off, float* rpn_class,
                    float* rpn_bbox, cudaStream_t st) {
  MRCNN_CHECK_CUDA(mrcnn_launch(rpn_post_kernel, dim3(grid_for((size_t)B * hw * apl, 256)), dim3(256), 0, st, head, ld, B, hw, apl, A, level_off, rpn_class, rpn_bbox));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_class_post(const float* head, int ld, int M, int NC, float* probs, float* bbox, cudaStream_t st) {
  MRCNN_CHECK_CUDA(mrcnn_launch(class_post_kernel, dim3(grid_for((size_t)M, 256)), dim3(256), 0, st, head, ld, M, NC, probs, bbox));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_mask_post(const float* logits, int ld, size_t M, int NC, float* out, cudaStream_t st) {
  MRCNN_CHECK_CUDA(mrcnn_launch(mask_post_kernel, dim3(grid_for(M * NC, 256)), dim3(256), 0, st, logits, ld, M, NC, out));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

int launch_mask_tile_flags(const float* det, int n_rois, int rows_per_roi, int tile_rows, int n_tiles, unsigned char* flags,
                           cudaStream_t st) {
  MRCNN_REQUIRE(det && flags && n_rois > 0 && rows_per_roi > 0 && tile_rows > 0 && n_tiles > 0, "mask_tile_flags: bad arguments");
  MRCNN_CHECK_CUDA(mrcnn_launch(mask_tile_flags_kernel, dim3((n_tiles + 255) / 256), dim3(256), 0, st, det, n_rois, rows_per_roi,
                                tile_rows, n_tiles, flags));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_mask_zero_padded(const float* det, int n_rois, size_t floats_per_roi, float* out, cudaStream_t st) {
  MRCNN_REQUIRE(det && out && n_rois > 0 && floats_per_roi % 4 == 0, "mask_zero_padded: bad arguments");
  MRCNN_CHECK_CUDA(mrcnn_launch(mask_zero_padded_kernel, dim3(n_rois), dim3(256), 0, st, det, floats_per_roi, out));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
