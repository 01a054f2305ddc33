// DetectionLayer / refine_detections_graph — one 1024-thread CTA per image:
//   argmax class, class-specific box decode (x BBOX_STD_DEV), clip to the image window,
//   background (and optional confidence) filter, per-class NMS in TF-1.13 pop order,
//   union ∩ keep (ascending), top-k by score, rows (y1,x1,y2,x2,class,score), zero pad.
// Replaces mrcnn/model.py:770-865 (refine_detections_graph), :868-909 (DetectionLayer.call),
// :3003-3017 (norm_boxes_graph) and the tf.map_fn while-loop over classes.
// Bit-exact contract: equals oracle/graph_layers.py detection_layer() for identical inputs.
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

constexpr int DET_THREADS = 1024;
constexpr int DET_MAX_N = 1024;
constexpr int DET_MAX_D = 256;

struct DetParams {
  const float* rois;    // [B,N,4]
  const float* probs;   // [B,N,NC]
  const float* deltas;  // [B,N,NC,4]
  const float* metas;   // [B,meta_size]
  int meta_size;
  int N, NC, D;
  float sd[4];
  float min_conf;  // 0 => branch skipped (truthiness, model.py:804)
  float thr;
  float* det;  // [B,D,6]
};

struct DetSmem {
  Box4 refined[DET_MAX_N];
  Box4 cboxes[DET_MAX_N];
  float score[DET_MAX_N];
  __align__(16) HeapEntry heap[DET_MAX_N + 2];   // 1-indexed
  unsigned long long sortbuf[DET_MAX_N];
  int cls[DET_MAX_N];
  uint16_t ixs[DET_MAX_N];
  uint16_t order[DET_MAX_N];
  uint16_t selected[DET_MAX_D];
  float4 kept_box[DET_MAX_D];
  float kept_area[DET_MAX_D];
  unsigned char keepmask[DET_MAX_N];
  unsigned char nmskeep[DET_MAX_N];
  NmsScratch sc;
  int warp_tot[32];
  int misc[4];
};

__device__ __forceinline__ int excl_scan_flag(bool flag, int* warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  int pre = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < nw; ++w) {
    int v = warp_tot[w];
    if (w < warp) off += v;
    tot += v;
  }
  __syncthreads();
  total = tot;
  return off + pre;
}

// descending bitonic sort of n_pad (power of two <= 1024) 64-bit keys, one element per thread pair
__device__ __forceinline__ void bitonic_desc(unsigned long long* buf, int n_pad) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_pad; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = buf[i], c = buf[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) {
            buf[i] = c;
            buf[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(DET_THREADS, 1) detection_kernel(DetParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DetSmem& s = *reinterpret_cast<DetSmem*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int N = p.N, NC = p.NC, D = p.D;

  // window in normalised coordinates; image_shape taken from the FIRST image (model.py:893-895)
  const float* m0 = p.metas;
  const float* mb = p.metas + (size_t)b * p.meta_size;
  const float ih = __fsub_rn(m0[4], 1.0f), iw = __fsub_rn(m0[5], 1.0f);
  const float wy1 = __fdiv_rn(__fsub_rn(mb[7], 0.0f), ih);
  const float wx1 = __fdiv_rn(__fsub_rn(mb[8], 0.0f), iw);
  const float wy2 = __fdiv_rn(__fsub_rn(mb[9], 1.0f), ih);
  const float wx2 = __fdiv_rn(__fsub_rn(mb[10], 1.0f), iw);

  // ---- 1. per-ROI class, score, refined box ---------------------------------------------------
  for (int i = tid; i < N; i += nt) {
    const float* pr = p.probs + ((size_t)b * N + i) * NC;
    int c = 0;
    float best = pr[0];
    for (int k = 1; k < NC; ++k) {
      float v = pr[k];
      if (v > best) {  // strict: first maximal index (tf.argmax)
        best = v;
        c = k;
      }
    }
    const float4 d = *reinterpret_cast<const float4*>(p.deltas + (((size_t)b * N + i) * NC + c) * 4);
    const float4 r = *reinterpret_cast<const float4*>(p.rois + ((size_t)b * N + i) * 4);
    Box4 bx = {r.x, r.y, r.z, r.w};
    bx = apply_box_deltas(bx, __fmul_rn(d.x, p.sd[0]), __fmul_rn(d.y, p.sd[1]),
                          __fmul_rn(d.z, p.sd[2]), __fmul_rn(d.w, p.sd[3]));
    s.refined[i] = clip_box(bx, wy1, wx1, wy2, wx2);
    s.cls[i] = c;
    s.score[i] = best;
    bool keep = c > 0;
    if (p.min_conf != 0.0f) keep = keep && (best >= p.min_conf);
    s.keepmask[i] = keep ? 1 : 0;
    s.nmskeep[i] = 0;
  }
  __syncthreads();

  // ---- 2. per-class NMS ------------------------------------------------------------------------
  for (int c = 1; c < NC; ++c) {
    // ascending list of kept ROIs of this class
    int n_c = 0;
    for (int base = 0; base < N; base += nt) {
      const int i = base + tid;
      const bool f = (i < N) && s.keepmask[i] && s.cls[i] == c;
      int tot;
      const int rank = n_c + excl_scan_flag(f, s.warp_tot, tot);
      if (f) s.ixs[rank] = (uint16_t)i;
      n_c += tot;
    }
    __syncthreads();
    if (n_c == 0) continue;
    int n_pad = 1;
    while (n_pad < n_c) n_pad <<= 1;
    for (int r = tid; r < n_pad; r += nt) {
      unsigned long long comp = 0ull;
      if (r < n_c) {
        const int i = s.ixs[r];
        s.cboxes[r] = s.refined[i];
        comp = ((unsigned long long)float_to_key(s.score[i]) << 32) | (uint32_t)(~(uint32_t)r);
      }
      s.sortbuf[r] = comp;
    }
    __syncthreads();
    bitonic_desc(s.sortbuf, n_pad);
    bool tie = false;
    for (int r = tid; r + 1 < n_c; r += nt)
      tie |= ((s.sortbuf[r] >> 32) == (s.sortbuf[r + 1] >> 32));
    const int any_tie = __syncthreads_or(tie ? 1 : 0);
    if (any_tie) {
      // TF pushes candidates in input (rank) order and pops through libstdc++'s heap
      for (int r = tid; r < n_c; r += nt) {
        HeapEntry e;
        e.score = s.score[s.ixs[r]];
        e.id = r;
        s.heap[r + 1] = e;
      }
      __syncthreads();
      if (tid == 0) heap_push_all_serial(s.heap, n_c);
    } else {
      for (int r = tid; r < n_c; r += nt)
        s.order[r] = (uint16_t)(~(uint32_t)(s.sortbuf[r] & 0xffffffffull));
    }
    __syncthreads();
    const int cnt = block_nms(s.cboxes, s.order, n_c, D, p.thr, s.kept_box, s.kept_area, s.selected, &s.sc, any_tie ? s.heap : nullptr);
    __syncthreads();
    for (int r = tid; r < cnt; r += nt) s.nmskeep[s.ixs[s.order[s.selected[r]]]] = 1;
    __syncthreads();
  }

  // ---- 3. keep ∩ nms_keep (ascending) -> top-k by score (ties: lower index first) ---------------
  int total = 0;
  for (int i = tid; i < DET_MAX_N; i += nt) s.sortbuf[i] = 0ull;
  __syncthreads();
  for (int base = 0; base < N; base += nt) {
    const int i = base + tid;
    const bool f = (i < N) && s.keepmask[i] && s.nmskeep[i];
    int tot;
    const int rank = total + excl_scan_flag(f, s.warp_tot, tot);
    if (f) s.sortbuf[rank] = ((unsigned long long)float_to_key(s.score[i]) << 32) | (uint32_t)(~(uint32_t)i);
    total += tot;
  }
  __syncthreads();
  int n_pad = 1;
  while (n_pad < total) n_pad <<= 1;
  bitonic_desc(s.sortbuf, n_pad);
  const int nout = total < D ? total : D;
  float* out = p.det + (size_t)b * D * 6;
  for (int r = tid; r < D; r += nt) {
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r < nout) {
      const int i = (int)(~(uint32_t)(s.sortbuf[r] & 0xffffffffull));
      const Box4 bx = s.refined[i];
      v[0] = bx.y1; v[1] = bx.x1; v[2] = bx.y2; v[3] = bx.x2;
      v[4] = (float)s.cls[i];
      v[5] = s.score[i];
    }
    for (int k = 0; k < 6; ++k) out[r * 6 + k] = v[k];
  }
}

}  // namespace

extern "C" int mrcnn_detection_layer(const float* rois, const float* mrcnn_class, const float* mrcnn_bbox,
                                     const float* image_metas, int meta_size, int batch, int num_rois,
                                     int num_classes, int max_instances, float min_confidence,
                                     float nms_threshold, const float* bbox_std_dev, float* detections,
                                     void* stream) {
  MRCNN_REQUIRE(rois && mrcnn_class && mrcnn_bbox && image_metas && detections && bbox_std_dev,
                "detection_layer: null pointer");
  MRCNN_REQUIRE(batch > 0, "detection_layer: empty batch");
  MRCNN_REQUIRE(num_rois >= 1 && num_rois <= DET_MAX_N, "detection_layer: num_rois=%d outside [1,%d]", num_rois, DET_MAX_N);
  MRCNN_REQUIRE(max_instances >= 1 && max_instances <= DET_MAX_D, "detection_layer: max_instances=%d outside [1,%d]", max_instances, DET_MAX_D);
  MRCNN_REQUIRE(num_classes >= 1 && meta_size >= 12, "detection_layer: bad num_classes/meta_size");
  DetParams p;
  p.rois = rois;
  p.probs = mrcnn_class;
  p.deltas = mrcnn_bbox;
  p.metas = image_metas;
  p.meta_size = meta_size;
  p.N = num_rois;
  p.NC = num_classes;
  p.D = max_instances;
  for (int i = 0; i < 4; ++i) p.sd[i] = bbox_std_dev[i];
  p.min_conf = min_confidence;
  p.thr = nms_threshold;
  p.det = detections;
  MRCNN_CHECK_CUDA(cudaFuncSetAttribute(detection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DetSmem)));
  MRCNN_CHECK_CUDA(mrcnn_launch(detection_kernel, dim3(batch), dim3(DET_THREADS), sizeof(DetSmem), static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
