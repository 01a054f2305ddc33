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

// A[(b*OH+oh)*OW+ow][k], k = (r*7+s)*3+c for k < 147, zero for 147 <= k < KPAD (=192)
__global__ void stem_im2col_kernel(const float* __restrict__ img, int B, int S, __nv_bfloat16* __restrict__ A) {
  constexpr int KPAD = 192, VEC = 8, VPR = KPAD / VEC;   // 24 vectors per row
  const int OH = S / 2, OW = S / 2;
  const size_t total = (size_t)B * OH * OW * VPR;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % VPR);
    size_t row = idx / VPR;
    const int ow = (int)(row % OW);
    const int oh = (int)((row / OW) % OH);
    const int b = (int)(row / ((size_t)OW * OH));
    float vals[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int k = v * VEC + j;
      float x = 0.f;
      if (k < 147) {
        const int tap = k / 3, c = k - tap * 3;
        const int r = tap / 7, s = tap - r * 7;
        const int ih = oh * 2 + r - 3, iw = ow * 2 + s - 3;
        if (ih >= 0 && ih < S && iw >= 0 && iw < S) x = __ldg(img + (((size_t)b * S + ih) * S + iw) * 3 + c);
      }
      vals[j] = x;
    }
    uint4 o = make_uint4(pack2(vals[0], vals[1]), pack2(vals[2], vals[3]), pack2(vals[4], vals[5]), pack2(vals[6], vals[7]));
    *reinterpret_cast<uint4*>(A + row * KPAD + (size_t)v * VEC) = o;
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
  const size_t total = M * NC;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / NC;
    const int k = (int)(idx - r * NC);
    const float v = logits[r * ld + k];
    out[idx] = 1.0f / (1.0f + expf(-v));
  }
}

int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int launch_stem_im2col(const float* img, int B, int S, __nv_bfloat16* A, cudaStream_t st) {
  const size_t total = (size_t)B * (S / 2) * (S / 2) * 24;
  stem_im2col_kernel<<<grid_for(total, 256), 256, 0, st>>>(img, B, S, A);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_maxpool3x3s2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st) {
  const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, B, H, W, C, y);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_subsample2(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st) {
  const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  subsample2_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, B, H, W, C, y);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_rpn_post(const float* head, int ld, int B, int hw, int apl, int A, int level_off, float* rpn_class,
                    float* rpn_bbox, cudaStream_t st) {
  rpn_post_kernel<<<grid_for((size_t)B * hw * apl, 256), 256, 0, st>>>(head, ld, B, hw, apl, A, level_off, rpn_class, rpn_bbox);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_class_post(const float* head, int ld, int M, int NC, float* probs, float* bbox, cudaStream_t st) {
  class_post_kernel<<<grid_for((size_t)M, 256), 256, 0, st>>>(head, ld, M, NC, probs, bbox);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
int launch_mask_post(const float* logits, int ld, size_t M, int NC, float* out, cudaStream_t st) {
  mask_post_kernel<<<grid_for(M * NC, 256), 256, 0, st>>>(logits, ld, M, NC, out);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
