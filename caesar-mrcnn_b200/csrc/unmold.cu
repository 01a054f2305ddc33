// unmold_detections on the device (a12): padded detections + 28x28 class masks -> pixel boxes,
// class ids, scores and full-frame boolean masks in the reference's [H, W, N] layout.
// Replaces mrcnn/model.py:2558-2621 (MaskRCNN.unmold_detections), mrcnn/utils.py:923-954
// (norm_boxes / denorm_boxes) and :629-645 (unmold_mask -> skimage<=0.15 resize, >= 0.5, paste),
// i.e. the host loop of per-detection skimage warps.
//   kernel 1 (one CTA per image): N = first class_id == 0; window-relative boxes; np.around to
//            int32; zero-area rows dropped (order preserved); per-mask min / max for the clip.
//   kernel 2 (one thread per output pixel): for every detection covering the pixel, float64
//            bilinear sample of the 28x28 mask with skimage semantics, clip, threshold; each pixel's
//            N-byte row is staged in shared memory and written coalesced (zero fill included).
// Bit-exact vs oracle/host_ops.py unmold_detections (float64 op order kept, no FMA contraction).
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

struct DetRec {
  int y1, x1, y2, x2;
  int cls, src;       // class id, original detection row
  double mn, mx;      // min / max of the selected 28x28 mask
  double rs, cs;      // MH / box height, MW / box width (skimage scale factors), hoisted out of the pixel loop
  double ro, co;      // 0.5 * rs - 0.5, 0.5 * cs - 0.5
};

struct UnmoldParams {
  const float* det;     // [B,D,6]
  const float* masks;   // [B,D,MH,MW,NC]
  int D, MH, MW, NC;
  int H0, W0, IH, IW;
  const int32_t* windows;  // [B,4]
  int32_t* rois;
  int32_t* class_ids;
  float* scores;
  int32_t* counts;
  DetRec* recs;         // [B,D]
  uint8_t* out;         // [B,H0,W0,D]
};

__global__ void unmold_boxes_kernel(UnmoldParams p) {
  pdl_prologue();
  extern __shared__ int s_int[];
  int* s_valid = s_int;            // [D]
  int* s_first0 = s_int + p.D;     // [1]
  const int b = blockIdx.x, tid = threadIdx.x, D = p.D;
  const float* det = p.det + (size_t)b * D * 6;
  if (tid == 0) *s_first0 = D;
  __syncthreads();
  for (int i = tid; i < D; i += blockDim.x)
    if (det[i * 6 + 4] == 0.f) atomicMin(s_first0, i);
  __syncthreads();
  const int N = *s_first0;
  // window in normalised coordinates: (window - [0,0,1,1]) / ([h,w,h,w]-1), float64 -> float32
  const int32_t* win = p.windows + b * 4;
  const float wy1 = (float)((double)(win[0] - 0) / (double)(p.IH - 1));
  const float wx1 = (float)((double)(win[1] - 0) / (double)(p.IW - 1));
  const float wy2 = (float)((double)(win[2] - 1) / (double)(p.IH - 1));
  const float wx2 = (float)((double)(win[3] - 1) / (double)(p.IW - 1));
  const float wh = __fsub_rn(wy2, wy1), ww = __fsub_rn(wx2, wx1);
  int y1 = 0, x1 = 0, y2 = 0, x2 = 0, cls = 0;
  float score = 0.f;
  bool valid = false;
  for (int i = tid; i < D; i += blockDim.x) s_valid[i] = 0;
  __syncthreads();
  const int i = tid;   // blockDim >= D
  if (i < N) {
    const float by1 = __fdiv_rn(__fsub_rn(det[i * 6 + 0], wy1), wh);
    const float bx1 = __fdiv_rn(__fsub_rn(det[i * 6 + 1], wx1), ww);
    const float by2 = __fdiv_rn(__fsub_rn(det[i * 6 + 2], wy1), wh);
    const float bx2 = __fdiv_rn(__fsub_rn(det[i * 6 + 3], wx1), ww);
    // denorm_boxes: np.around(boxes * [h-1,w-1,h-1,w-1] + [0,0,1,1]) in float64 -> int32
    y1 = (int)rint(__dadd_rn(__dmul_rn((double)by1, (double)(p.H0 - 1)), 0.0));
    x1 = (int)rint(__dadd_rn(__dmul_rn((double)bx1, (double)(p.W0 - 1)), 0.0));
    y2 = (int)rint(__dadd_rn(__dmul_rn((double)by2, (double)(p.H0 - 1)), 1.0));
    x2 = (int)rint(__dadd_rn(__dmul_rn((double)bx2, (double)(p.W0 - 1)), 1.0));
    cls = (int)det[i * 6 + 4];
    score = det[i * 6 + 5];
    valid = ((long long)(y2 - y1) * (long long)(x2 - x1)) > 0;
    s_valid[i] = valid ? 1 : 0;
  }
  __syncthreads();
  int slot = 0, total = 0;
  for (int k = 0; k < D; ++k) {
    const int v = s_valid[k];
    if (k < i) slot += v;
    total += v;
  }
  if (tid == 0) p.counts[b] = total;
  if (i < D) {   // clear the padded tail, fill the compacted head
    if (i >= total) {
      int32_t* r = p.rois + ((size_t)b * D + i) * 4;
      r[0] = r[1] = r[2] = r[3] = 0;
      p.class_ids[(size_t)b * D + i] = 0;
      p.scores[(size_t)b * D + i] = 0.f;
    }
  }
  int* s_slot = s_first0 + 1;      // [D] compacted slot of detection i, -1 = dropped
  int* s_cls = s_slot + D;         // [D]
  if (i < D) {
    s_slot[i] = valid ? slot : -1;
    s_cls[i] = cls;
  }
  if (valid) {
    int32_t* r = p.rois + ((size_t)b * D + slot) * 4;
    r[0] = y1; r[1] = x1; r[2] = y2; r[3] = x2;
    p.class_ids[(size_t)b * D + slot] = cls;
    p.scores[(size_t)b * D + slot] = score;
    DetRec* rec = p.recs + (size_t)b * D + slot;
    rec->y1 = y1; rec->x1 = x1; rec->y2 = y2; rec->x2 = x2; rec->cls = cls; rec->src = i;
    const double rs = (double)p.MH / (double)(y2 - y1), cs = (double)p.MW / (double)(x2 - x1);
    rec->rs = rs;
    rec->cs = cs;
    rec->ro = __dsub_rn(__dmul_rn(0.5, rs), 0.5);
    rec->co = __dsub_rn(__dmul_rn(0.5, cs), 0.5);
  }
  __syncthreads();
  // min / max of each kept detection's class mask (skimage clip range): one warp per detection, lanes stride
  // over the MH*MW samples (min / max are order-independent, so the lane split is exact)
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int d = warp; d < D; d += nwarps) {
    const int sl = s_slot[d];
    if (sl < 0) continue;
    const float* m = p.masks + (((size_t)b * D + d) * p.MH * p.MW) * p.NC + s_cls[d];
    float mn = INFINITY, mx = -INFINITY;
    for (int k = lane; k < p.MH * p.MW; k += 32) {
      const float v = __ldg(m + (size_t)k * p.NC);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      DetRec* rec = p.recs + (size_t)b * D + sl;
      rec->mn = (double)mn;
      rec->mx = (double)mx;
    }
  }
}

constexpr int PAINT_THREADS = 128;

__global__ void __launch_bounds__(PAINT_THREADS) unmold_paint_kernel(UnmoldParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int D = p.D;
  DetRec* s_rec = reinterpret_cast<DetRec*>(s_raw);                         // [D]
  unsigned char* s_row = s_raw + (((size_t)D * sizeof(DetRec) + 15) & ~(size_t)15);  // [PAINT_THREADS * D]
  const int b = blockIdx.y, tid = threadIdx.x;
  const int cnt = p.counts[b];
  for (int i = tid; i < cnt; i += blockDim.x) s_rec[i] = p.recs[(size_t)b * D + i];
  const size_t npx = (size_t)p.H0 * p.W0;
  const size_t p0 = (size_t)blockIdx.x * PAINT_THREADS;
  // zero the staging rows (D bytes per pixel)
  const int stage_words = (PAINT_THREADS * D + 3) / 4;
  for (int i = tid; i < stage_words; i += blockDim.x) reinterpret_cast<uint32_t*>(s_row)[i] = 0u;
  // detections whose box intersects this CTA's pixel span (PAINT_THREADS consecutive pixels in raster order): most
  // boxes cover a small part of the frame, so the per-pixel loop below runs over a short list instead of all cnt
  int* s_list = reinterpret_cast<int*>(s_row + (size_t)PAINT_THREADS * D);          // [D]
  int* s_nlist = s_list + D;
  if (tid == 0) *s_nlist = 0;
  __syncthreads();
  {
    const size_t plast = min(p0 + PAINT_THREADS, npx) - 1;
    const int ya = (int)(p0 / p.W0), yb = (int)(plast / p.W0);
    const int xa = (int)(p0 % p.W0), xb = (int)(plast % p.W0);
    for (int k = tid; k < cnt; k += blockDim.x) {
      const DetRec& r = s_rec[k];
      bool hit = r.y1 <= yb && r.y2 > ya;
      if (hit && ya == yb) hit = r.x1 <= xb && r.x2 > xa;       // span inside one image row: clip in x as well
      if (hit) s_list[atomicAdd(s_nlist, 1)] = k;
    }
  }
  __syncthreads();
  const int nlist = *s_nlist;
  const size_t pix = p0 + tid;
  if (pix < npx) {
    const int y = (int)(pix / p.W0), x = (int)(pix % p.W0);
    unsigned char* row = s_row + (size_t)tid * D;
    for (int li = 0; li < nlist; ++li) {
      const int k = s_list[li];                                  // order is irrelevant: each k writes its own byte
      const DetRec& r = s_rec[k];
      if (y < r.y1 || y >= r.y2 || x < r.x1 || x >= r.x2) continue;
      const double rr = __dadd_rn(__dmul_rn(r.rs, (double)(y - r.y1)), r.ro);
      const double cc = __dadd_rn(__dmul_rn(r.cs, (double)(x - r.x1)), r.co);
      const double fr = floor(rr), fc = floor(cc);
      const int r0 = (int)fr, r1 = (int)ceil(rr), c0 = (int)fc, c1 = (int)ceil(cc);
      const double dr = __dsub_rn(rr, fr), dc = __dsub_rn(cc, fc);
      const double wr = __dsub_rn(1.0, dr), wc = __dsub_rn(1.0, dc);
      const float* m = p.masks + (((size_t)b * D + r.src) * p.MH * p.MW) * p.NC + r.cls;
      auto px = [&](int ri, int ci) -> double {
        if (ri < 0 || ri >= p.MH || ci < 0 || ci >= p.MW) return 0.0;
        return (double)__ldg(m + ((size_t)ri * p.MW + ci) * p.NC);
      };
      const double t = __dadd_rn(__dmul_rn(wc, px(r0, c0)), __dmul_rn(dc, px(r0, c1)));
      const double bo = __dadd_rn(__dmul_rn(wc, px(r1, c0)), __dmul_rn(dc, px(r1, c1)));
      double v = __dadd_rn(__dmul_rn(wr, t), __dmul_rn(dr, bo));
      const bool preserve_cval = !(r.mn <= 0.0 && 0.0 <= r.mx);
      if (!(preserve_cval && v == 0.0)) v = fmin(fmax(v, r.mn), r.mx);
      row[k] = v >= 0.5 ? 1 : 0;
    }
  }
  __syncthreads();
  // coalesced write of the CTA's contiguous [pixels, D] byte block
  const size_t nvalid = (p0 + PAINT_THREADS <= npx) ? PAINT_THREADS : (npx > p0 ? npx - p0 : 0);
  const size_t nbytes = nvalid * D;
  unsigned char* dst = p.out + ((size_t)b * npx + p0) * D;
  if (((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const size_t nv = nbytes / 16;
    for (size_t i = tid; i < nv; i += blockDim.x)
      __stcs(reinterpret_cast<uint4*>(dst) + i, reinterpret_cast<const uint4*>(s_row)[i]);
    for (size_t i = nv * 16 + tid; i < nbytes; i += blockDim.x) dst[i] = s_row[i];
  } else {
    for (size_t i = tid; i < nbytes; i += blockDim.x) dst[i] = s_row[i];
  }
}


// Same per-pixel arithmetic as unmold_paint_kernel, but the result leaves as PIXEL-MAJOR BITS: out[b][pixel][DW] uint32,
// bit k of word w = detection 32*w + k of image b (compacted order), DW = mask_bits_words(D) in {1,2,4,8}.  One thread per
// pixel keeps its DW words in registers and stores them as one aligned vector: 16 bytes per pixel for D = 100 instead
// of 100, which is what crosses PCIe on the host-result path (the [H,W,N] bool arrays of the reference contract are
// expanded from these bits on the host, host_expand.cu).
template <int DW>
__global__ void __launch_bounds__(PAINT_THREADS) unmold_paint_bits_kernel(UnmoldParams p, uint32_t* __restrict__ out_bits) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int D = p.D;
  DetRec* s_rec = reinterpret_cast<DetRec*>(s_raw);                                  // [D]
  int* s_list = reinterpret_cast<int*>(s_raw + (((size_t)D * sizeof(DetRec) + 15) & ~(size_t)15));   // [D]
  int* s_nlist = s_list + D;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int cnt = p.counts[b];
  for (int i = tid; i < cnt; i += blockDim.x) s_rec[i] = p.recs[(size_t)b * D + i];
  const size_t npx = (size_t)p.H0 * p.W0;
  const size_t p0 = (size_t)blockIdx.x * PAINT_THREADS;
  if (tid == 0) *s_nlist = 0;
  __syncthreads();
  {
    const size_t plast = min(p0 + PAINT_THREADS, npx) - 1;
    const int ya = (int)(p0 / p.W0), yb = (int)(plast / p.W0);
    const int xa = (int)(p0 % p.W0), xb = (int)(plast % p.W0);
    for (int k = tid; k < cnt; k += blockDim.x) {
      const DetRec& r = s_rec[k];
      bool hit = r.y1 <= yb && r.y2 > ya;
      if (hit && ya == yb) hit = r.x1 <= xb && r.x2 > xa;
      if (hit) s_list[atomicAdd(s_nlist, 1)] = k;
    }
  }
  __syncthreads();
  const int nlist = *s_nlist;
  const size_t pix = p0 + tid;
  if (pix >= npx) return;
  uint32_t words[DW];
#pragma unroll
  for (int w = 0; w < DW; ++w) words[w] = 0u;
  const int y = (int)(pix / p.W0), x = (int)(pix % p.W0);
  for (int li = 0; li < nlist; ++li) {
    const int k = s_list[li];
    const DetRec& r = s_rec[k];
    if (y < r.y1 || y >= r.y2 || x < r.x1 || x >= r.x2) continue;
    const double rr = __dadd_rn(__dmul_rn(r.rs, (double)(y - r.y1)), r.ro);
    const double cc = __dadd_rn(__dmul_rn(r.cs, (double)(x - r.x1)), r.co);
    const double fr = floor(rr), fc = floor(cc);
    const int r0 = (int)fr, r1 = (int)ceil(rr), c0 = (int)fc, c1 = (int)ceil(cc);
    const double dr = __dsub_rn(rr, fr), dc = __dsub_rn(cc, fc);
    const double wr = __dsub_rn(1.0, dr), wc = __dsub_rn(1.0, dc);
    const float* m = p.masks + (((size_t)b * D + r.src) * p.MH * p.MW) * p.NC + r.cls;
    auto px = [&](int ri, int ci) -> double {
      if (ri < 0 || ri >= p.MH || ci < 0 || ci >= p.MW) return 0.0;
      return (double)__ldg(m + ((size_t)ri * p.MW + ci) * p.NC);
    };
    const double t = __dadd_rn(__dmul_rn(wc, px(r0, c0)), __dmul_rn(dc, px(r0, c1)));
    const double bo = __dadd_rn(__dmul_rn(wc, px(r1, c0)), __dmul_rn(dc, px(r1, c1)));
    double v = __dadd_rn(__dmul_rn(wr, t), __dmul_rn(dr, bo));
    const bool preserve_cval = !(r.mn <= 0.0 && 0.0 <= r.mx);
    if (!(preserve_cval && v == 0.0)) v = fmin(fmax(v, r.mn), r.mx);
    if (v >= 0.5) {
#pragma unroll
      for (int w = 0; w < DW; ++w)
        if ((k >> 5) == w) words[w] |= 1u << (k & 31);
    }
  }
  uint32_t* dst = out_bits + ((size_t)b * npx + pix) * DW;
  if constexpr (DW == 1) {
    __stcs(dst, words[0]);
  } else if constexpr (DW == 2) {
    __stcs(reinterpret_cast<uint2*>(dst), make_uint2(words[0], words[1]));
  } else {
#pragma unroll
    for (int w = 0; w < DW; w += 4)
      __stcs(reinterpret_cast<uint4*>(dst + w), make_uint4(words[w], words[w + 1], words[w + 2], words[w + 3]));
  }
}

}  // namespace

extern "C" size_t mrcnn_unmold_workspace_bytes(int batch, int max_instances) {
  return (size_t)batch * (size_t)max_instances * sizeof(DetRec);
}

extern "C" int mrcnn_unmold_detections(const float* detections, const float* mrcnn_mask, int batch, int max_instances,
                                       int mask_h, int mask_w, int num_classes, const int* orig_hw, const int* image_hw,
                                       const int32_t* windows, int32_t* rois, int32_t* class_ids, float* scores,
                                       int32_t* counts, uint8_t* masks, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  MRCNN_REQUIRE(detections && mrcnn_mask && orig_hw && image_hw && windows && rois && class_ids && scores && counts && masks,
                "unmold_detections: null pointer");
  MRCNN_REQUIRE(batch > 0 && max_instances > 0 && max_instances <= 256, "unmold_detections: batch/max_instances out of range");
  MRCNN_REQUIRE(mask_h > 0 && mask_w > 0 && num_classes > 0, "unmold_detections: bad mask shape");
  MRCNN_REQUIRE(orig_hw[0] > 1 && orig_hw[1] > 1 && image_hw[0] > 1 && image_hw[1] > 1, "unmold_detections: bad image size");
  MRCNN_REQUIRE(workspace && workspace_bytes >= mrcnn_unmold_workspace_bytes(batch, max_instances),
                "unmold_detections: workspace too small");
  UnmoldParams p;
  p.det = detections; p.masks = mrcnn_mask; p.D = max_instances; p.MH = mask_h; p.MW = mask_w; p.NC = num_classes;
  p.H0 = orig_hw[0]; p.W0 = orig_hw[1]; p.IH = image_hw[0]; p.IW = image_hw[1];
  p.windows = windows; p.rois = rois; p.class_ids = class_ids; p.scores = scores; p.counts = counts;
  p.recs = static_cast<DetRec*>(workspace); p.out = masks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 1024;        // >= max_instances (<= 256): thread i = detection i, then one warp per detection
  MRCNN_CHECK_CUDA(mrcnn_launch(unmold_boxes_kernel, dim3(batch), dim3(threads), (3 * max_instances + 1) * sizeof(int), st, p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  const size_t npx = (size_t)p.H0 * p.W0;
  const size_t smem = (((size_t)max_instances * sizeof(DetRec) + 15) & ~(size_t)15) + (size_t)PAINT_THREADS * max_instances +
                      (size_t)(max_instances + 1) * sizeof(int) + 16;
  MRCNN_CHECK_CUDA(cudaFuncSetAttribute(unmold_paint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRCNN_CHECK_CUDA(mrcnn_launch(unmold_paint_kernel, dim3(dim3((unsigned)((npx + PAINT_THREADS - 1) / PAINT_THREADS), batch)), dim3(PAINT_THREADS), smem, st, p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(2);
  return MRCNN_OK;
}

namespace {
// bits [n_images][npx][dw] -> out [n_images][npx * D] bytes holding, per image, the dense C-order [npx, counts[i]] array
// (the reference's [H, W, N] bool layout) in the first npx * counts[i] bytes of its slot.  One thread per (pixel, 32-bit
// word): 4 bits become 4 bytes with one multiply ((v & 15) * 0x00204081 & 0x01010101: the 16 partial products land on 16
// different bit positions, so nothing carries); word stores when the row length is a multiple of 4, byte stores otherwise.
__global__ void mask_bits_expand_kernel(const uint32_t* __restrict__ bits, const int32_t* __restrict__ counts, uint32_t npx, int dw,
                                        int D, uint8_t* __restrict__ out) {
  const int img = blockIdx.y;
  const int n = counts[img];
  if (n <= 0) return;
  const uint32_t* ib = bits + (size_t)img * npx * dw;
  uint8_t* ob = out + (size_t)img * npx * D;
  const uint32_t items = npx * (uint32_t)dw;
  const bool aligned = (n & 3) == 0;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < items; t += gridDim.x * blockDim.x) {
    const uint32_t p = t / (uint32_t)dw, j = t - p * (uint32_t)dw;       // dw is 1, 2, 4 or 8
    const int k0 = (int)j * 32;
    if (k0 >= n) continue;
    const int cnt = n - k0 < 32 ? n - k0 : 32;
    const uint32_t x = ib[t];
    uint8_t* dst = ob + (size_t)p * n + k0;
    if (aligned) {
      for (int q = 0; q * 4 < cnt; ++q)
        *reinterpret_cast<uint32_t*>(dst + 4 * q) = (((x >> (4 * q)) & 15u) * 0x00204081u) & 0x01010101u;
    } else {
      for (int k = 0; k < cnt; ++k) dst[k] = (uint8_t)((x >> k) & 1u);
    }
  }
}
}  // namespace

extern "C" int mrcnn_mask_bits_expand_device(const uint32_t* mask_bits, const int32_t* counts, int n_images,
                                             int64_t pixels_per_image, int max_instances, uint8_t* dense, void* stream) {
  MRCNN_REQUIRE(mask_bits && counts && dense && n_images > 0 && pixels_per_image > 0, "mask_bits_expand_device: bad arguments");
  MRCNN_REQUIRE(max_instances > 0 && ((size_t)pixels_per_image * max_instances) % 4 == 0,
                "mask_bits_expand_device: image slots (pixels * max_instances bytes) must be 4-byte multiples");
  MRCNN_REQUIRE(n_images <= 65535, "mask_bits_expand_device: too many images");
  const int dw = mrcnn_mask_bits_words(max_instances);
  MRCNN_REQUIRE((unsigned long long)pixels_per_image * dw < (1ull << 31), "mask_bits_expand_device: image too large");
  const size_t items = (size_t)pixels_per_image * dw;
  const unsigned gx = (unsigned)((items + 255) / 256 < 592 ? (items + 255) / 256 : 592);
  mask_bits_expand_kernel<<<dim3(gx, n_images), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mask_bits, counts, (uint32_t)pixels_per_image, dw, max_instances, dense);
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_mask_bits_words(int max_instances) {
  const int need = (max_instances + 31) / 32;
  int dw = 1;
  while (dw < need) dw <<= 1;
  return dw;
}

extern "C" int mrcnn_unmold_detections_bits(const float* detections, const float* mrcnn_mask, int batch, int max_instances,
                                            int mask_h, int mask_w, int num_classes, const int* orig_hw, const int* image_hw,
                                            const int32_t* windows, int32_t* rois, int32_t* class_ids, float* scores,
                                            int32_t* counts, uint32_t* mask_bits, void* workspace, size_t workspace_bytes,
                                            void* stream) {
  MRCNN_REQUIRE(detections && mrcnn_mask && orig_hw && image_hw && windows && rois && class_ids && scores && counts && mask_bits,
                "unmold_detections_bits: null pointer");
  MRCNN_REQUIRE(batch > 0 && max_instances > 0 && max_instances <= 256, "unmold_detections_bits: batch/max_instances out of range");
  MRCNN_REQUIRE(mask_h > 0 && mask_w > 0 && num_classes > 0, "unmold_detections_bits: bad mask shape");
  MRCNN_REQUIRE(orig_hw[0] > 1 && orig_hw[1] > 1 && image_hw[0] > 1 && image_hw[1] > 1, "unmold_detections_bits: bad image size");
  MRCNN_REQUIRE(workspace && workspace_bytes >= mrcnn_unmold_workspace_bytes(batch, max_instances),
                "unmold_detections_bits: workspace too small");
  UnmoldParams p;
  p.det = detections; p.masks = mrcnn_mask; p.D = max_instances; p.MH = mask_h; p.MW = mask_w; p.NC = num_classes;
  p.H0 = orig_hw[0]; p.W0 = orig_hw[1]; p.IH = image_hw[0]; p.IW = image_hw[1];
  p.windows = windows; p.rois = rois; p.class_ids = class_ids; p.scores = scores; p.counts = counts;
  p.recs = static_cast<DetRec*>(workspace); p.out = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MRCNN_CHECK_CUDA(mrcnn_launch(unmold_boxes_kernel, dim3(batch), dim3(1024), (3 * max_instances + 1) * sizeof(int), st, p));
  const size_t npx = (size_t)p.H0 * p.W0;
  const size_t smem = (((size_t)max_instances * sizeof(DetRec) + 15) & ~(size_t)15) + (size_t)(max_instances + 1) * sizeof(int) + 16;
  const dim3 grid((unsigned)((npx + PAINT_THREADS - 1) / PAINT_THREADS), batch);
  switch (mrcnn_mask_bits_words(max_instances)) {
    case 1: MRCNN_CHECK_CUDA(mrcnn_launch(unmold_paint_bits_kernel<1>, grid, dim3(PAINT_THREADS), smem, st, p, mask_bits)); break;
    case 2: MRCNN_CHECK_CUDA(mrcnn_launch(unmold_paint_bits_kernel<2>, grid, dim3(PAINT_THREADS), smem, st, p, mask_bits)); break;
    case 4: MRCNN_CHECK_CUDA(mrcnn_launch(unmold_paint_bits_kernel<4>, grid, dim3(PAINT_THREADS), smem, st, p, mask_bits)); break;
    default: MRCNN_CHECK_CUDA(mrcnn_launch(unmold_paint_bits_kernel<8>, grid, dim3(PAINT_THREADS), smem, st, p, mask_bits)); break;
  }
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(2);
  return MRCNN_OK;
}
