// Train-mode graph layers that are not dense contractions (BASELINE.json configs[4]):
//   detection_targets_kernel   <- mrcnn/model.py:570-683 detection_targets_graph (+ overlaps_graph :540-567,
//                                 utils.box_refinement_graph utils.py:249-272, tf.image.crop_and_resize of the GT masks)
//   roialign_backward_kernel   <- gradient of PyramidROIAlign (mrcnn/model.py:428-534): scatter-add of the pooled
//                                 gradients into float32 pyramid gradients with the forward's bilinear weights
//   sgd_norm / sgd_apply       <- keras.optimizers.SGD(lr, momentum, clipnorm) + the L2 weight regulariser of
//                                 MaskRCNN.compile (mrcnn/model.py:2259-2297), over ONE flat parameter buffer
// Arithmetic conventions follow the detect-path kernels: float32, one IEEE rounding per reference op, log = "double log,
// one rounding to float".  tf.random.shuffle is replaced by a documented counter hash (shuffle_key below): the order of a
// shuffle is not reproducible across TF builds anyway; the oracle (oracle/train_ops.py) uses the same keys.
#include <float.h>
#include <stdlib.h>
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

// ------------------------------------------------------------------------------------------------------------------
// DetectionTargetLayer
// ------------------------------------------------------------------------------------------------------------------
constexpr int DT_THREADS = 512;
constexpr int DT_MAX_N = 2048;     // proposals per image
constexpr int DT_MAX_G = 512;      // ground-truth instances per image

struct DetTargetParams {
  const float* proposals;        // [B,N,4] normalised, zero rows = padding
  const int32_t* gt_class_ids;   // [B,G]   (<0: crowd, 0: padding)
  const float* gt_boxes;         // [B,G,4] normalised
  const uint8_t* gt_masks;       // [B,MH,MW,G] 0/1
  int N, G, MH, MW, T, mask_h, mask_w, use_mini_mask, positive_count;
  float inv_ratio;               // float32(1 / ROI_POSITIVE_RATIO)
  float std_dev[4];
  unsigned long long seed;
  const unsigned long long* seed_device;   // optional: added to seed (a replayed CUDA graph advances it on the device)
  float* rois;                   // [B,T,4]
  int32_t* target_class_ids;     // [B,T]
  float* target_bbox;            // [B,T,4]
  float* target_mask;            // [B,T,mask_h,mask_w]
  int32_t* counts;               // [B,2] (positives, negatives) or null
};

// murmur3 finaliser over (seed, image, stream, index): the "random" sort key of a shuffle
__device__ __host__ inline uint32_t shuffle_key(unsigned long long seed, uint32_t image, uint32_t stream, uint32_t index) {
  uint32_t x = (uint32_t)seed ^ (uint32_t)(seed >> 32) * 0x9E3779B9u;
  x ^= image * 0x85EBCA6Bu + 0x27D4EB2Fu;
  x ^= stream * 0xC2B2AE35u;
  x ^= index * 0x165667B1u + 0x9E3779B9u;
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

// overlaps_graph (mrcnn/model.py:540-567), float32, one rounding per op
__device__ __forceinline__ float iou_tf(const float4 a, const float4 b) {   // (y1,x1,y2,x2)
  const float y1 = fmaxf(a.x, b.x), x1 = fmaxf(a.y, b.y), y2 = fminf(a.z, b.z), x2 = fminf(a.w, b.w);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f));
  const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
  return __fdiv_rn(inter, uni);
}

__device__ inline void bitonic_sort_asc(unsigned long long* buf, int npad) {
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = buf[i], c = buf[ixj];
          const bool up = (i & k) == 0;
          if (up ? (a > c) : (a < c)) {
            buf[i] = c;
            buf[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(DT_THREADS) detection_targets_kernel(DetTargetParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem[];
  float4* pbox = reinterpret_cast<float4*>(smem);                              // [N] trimmed proposals
  unsigned long long* pos = reinterpret_cast<unsigned long long*>(pbox + DT_MAX_N);   // [N pad]
  unsigned long long* neg = pos + DT_MAX_N;                                    // [N pad]
  float4* gbox = reinterpret_cast<float4*>(neg + DT_MAX_N);                    // [G] non-crowd GT boxes
  float4* cbox = gbox + DT_MAX_G;                                              // [G] crowd boxes
  int* gcls = reinterpret_cast<int*>(cbox + DT_MAX_G);                         // [G]
  int* gsrc = gcls + DT_MAX_G;                                                 // [G] index into the padded GT arrays
  short* assign = reinterpret_cast<short*>(gsrc + DT_MAX_G);                   // [N] argmax GT of each trimmed proposal
  __shared__ int s_np, s_ng, s_nc, s_npos, s_nneg, s_warp[DT_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const unsigned long long seed = p.seed + (p.seed_device ? *p.seed_device : 0ull);
  const float4* props = reinterpret_cast<const float4*>(p.proposals) + (size_t)b * p.N;
  const float4* gtb = reinterpret_cast<const float4*>(p.gt_boxes) + (size_t)b * p.G;
  const int32_t* gtc = p.gt_class_ids + (size_t)b * p.G;

  // -- GT: drop zero boxes (trim_zeros_graph: sum |coords| == 0), split crowd / non-crowd, keep order ---------------
  if (tid == 0) {
    int ng = 0, nc = 0;
    for (int g = 0; g < p.G; ++g) {
      const float4 bx = gtb[g];
      const float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(bx.x), fabsf(bx.y)), fabsf(bx.z)), fabsf(bx.w));
      if (!(s != 0.f)) continue;
      const int c = gtc[g];
      if (c < 0) {
        cbox[nc++] = bx;
      } else if (c > 0) {
        gbox[ng] = bx;
        gcls[ng] = c;
        gsrc[ng] = g;
        ++ng;
      }
    }
    s_ng = ng;
    s_nc = nc;
    s_npos = 0;
    s_nneg = 0;
  }
  // -- proposals: drop zero rows, keep order (block-wide ordered compaction) -------------------------------------------
  int base = 0;
  for (int start = 0; start < p.N; start += DT_THREADS) {
    const int i = start + tid;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    bool keep = false;
    if (i < p.N) {
      bx = props[i];
      keep = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(bx.x), fabsf(bx.y)), fabsf(bx.z)), fabsf(bx.w)) != 0.f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
    for (int w = 0; w < DT_THREADS / 32; ++w) {
      if (w < (tid >> 5)) off += s_warp[w];
      tot += s_warp[w];
    }
    if (keep) pbox[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = bx;
    base += tot;
    __syncthreads();
  }
  if (tid == 0) s_np = base;
  __syncthreads();
  const int np = s_np, ng = s_ng, nc = s_nc;

  // -- overlaps, positive / negative candidate lists with shuffle keys ---------------------------------------------------
  for (int i = tid; i < np; i += DT_THREADS) {
    const float4 bx = pbox[i];
    float best = -FLT_MAX * 2.f;      // -inf: tf.reduce_max of an empty row
    int arg = 0;
    for (int g = 0; g < ng; ++g) {
      const float v = iou_tf(bx, gbox[g]);
      if (v > best) {                  // first maximum (tf.argmax)
        best = v;
        arg = g;
      }
    }
    float crowd = -FLT_MAX * 2.f;
    for (int g = 0; g < nc; ++g) crowd = fmaxf(crowd, iou_tf(bx, cbox[g]));
    assign[i] = (short)arg;
    if (best >= 0.5f) {
      pos[atomicAdd(&s_npos, 1)] = ((unsigned long long)shuffle_key(seed, b, 0, i) << 32) | (unsigned)i;
    } else if (best < 0.5f && crowd < 0.001f) {
      neg[atomicAdd(&s_nneg, 1)] = ((unsigned long long)shuffle_key(seed, b, 1, i) << 32) | (unsigned)i;
    }
  }
  __syncthreads();
  const int npos = s_npos, nneg = s_nneg;
  int ppad = 1, qpad = 1;
  while (ppad < npos) ppad <<= 1;
  while (qpad < nneg) qpad <<= 1;
  for (int i = npos + tid; i < ppad; i += DT_THREADS) pos[i] = ~0ull;
  for (int i = nneg + tid; i < qpad; i += DT_THREADS) neg[i] = ~0ull;
  __syncthreads();
  bitonic_sort_asc(pos, ppad);      // ascending (key, index): the shuffled order
  bitonic_sort_asc(neg, qpad);
  const int pc = min(npos, p.positive_count);
  // negative_count = int32(r * float32(positive_count)) - positive_count  (mrcnn/model.py:644-646)
  const int want_neg = (int)__fmul_rn(p.inv_ratio, (float)pc) - pc;
  const int ncnt = min(nneg, max(want_neg, 0));
  const int T = p.T;
  if (tid == 0 && p.counts) {
    p.counts[2 * b] = min(pc, T);
    p.counts[2 * b + 1] = min(ncnt, max(T - pc, 0));
  }

  // -- rois / class ids / deltas, zero padded to T (rows beyond T cannot exist: pc + ncnt <= T by construction) ----------
  float4* rois = reinterpret_cast<float4*>(p.rois) + (size_t)b * T;
  float4* tbox = reinterpret_cast<float4*>(p.target_bbox) + (size_t)b * T;
  int32_t* tcls = p.target_class_ids + (size_t)b * T;
  for (int t = tid; t < T; t += DT_THREADS) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f), d = r;
    int c = 0;
    if (t < pc) {
      const int i = (int)(pos[t] & 0xffffffffu);
      r = pbox[i];
      const int g = assign[i];
      const float4 gt = gbox[g];
      c = gcls[g];
      // utils.box_refinement_graph (utils.py:249-272), then /= BBOX_STD_DEV
      const float h = __fsub_rn(r.z, r.x), w = __fsub_rn(r.w, r.y);
      const float cy = __fadd_rn(r.x, __fmul_rn(0.5f, h)), cx = __fadd_rn(r.y, __fmul_rn(0.5f, w));
      const float gh = __fsub_rn(gt.z, gt.x), gw = __fsub_rn(gt.w, gt.y);
      const float gcy = __fadd_rn(gt.x, __fmul_rn(0.5f, gh)), gcx = __fadd_rn(gt.y, __fmul_rn(0.5f, gw));
      d.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, cy), h), p.std_dev[0]);
      d.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, cx), w), p.std_dev[1]);
      d.z = __fdiv_rn((float)log((double)__fdiv_rn(gh, h)), p.std_dev[2]);
      d.w = __fdiv_rn((float)log((double)__fdiv_rn(gw, w)), p.std_dev[3]);
    } else if (t < pc + ncnt) {
      r = pbox[(int)(neg[t - pc] & 0xffffffffu)];
    }
    rois[t] = r;
    tbox[t] = d;
    tcls[t] = c;
  }

  // -- mask targets: crop_and_resize of the assigned GT mask to mask_h x mask_w, rounded (half to even) -------------------
  const int mh = p.mask_h, mw = p.mask_w, MH = p.MH, MW = p.MW, G = p.G;
  const uint8_t* masks = p.gt_masks + (size_t)b * MH * MW * G;
  float* tm = p.target_mask + (size_t)b * T * mh * mw;
  const int per = mh * mw;
  for (int e = tid; e < T * per; e += DT_THREADS) {
    const int t = e / per, r = e - t * per, iy = r / mw, ix = r - iy * mw;
    float v = 0.f;
    if (t < pc) {
      const int i = (int)(pos[t] & 0xffffffffu);
      const int g = assign[i];
      float4 bx = pbox[i];
      if (p.use_mini_mask) {        // ROI in the normalised frame of its GT box (mrcnn/model.py:648-659)
        const float4 gt = gbox[g];
        const float gh = __fsub_rn(gt.z, gt.x), gw = __fsub_rn(gt.w, gt.y);
        bx = make_float4(__fdiv_rn(__fsub_rn(bx.x, gt.x), gh), __fdiv_rn(__fsub_rn(bx.y, gt.y), gw),
                         __fdiv_rn(__fsub_rn(bx.z, gt.x), gh), __fdiv_rn(__fsub_rn(bx.w, gt.y), gw));
      }
      const float Hm1 = (float)(MH - 1), Wm1 = (float)(MW - 1);
      const float sy = mh > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(bx.z, bx.x), Hm1), (float)(mh - 1)) : 0.f;
      const float sx = mw > 1 ? __fdiv_rn(__fmul_rn(__fsub_rn(bx.w, bx.y), Wm1), (float)(mw - 1)) : 0.f;
      const float in_y = mh > 1 ? __fadd_rn(__fmul_rn(bx.x, Hm1), __fmul_rn((float)iy, sy))
                                : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(bx.x, bx.z)), Hm1);
      const float in_x = mw > 1 ? __fadd_rn(__fmul_rn(bx.y, Wm1), __fmul_rn((float)ix, sx))
                                : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(bx.y, bx.w)), Wm1);
      if (in_y >= 0.f && in_y <= Hm1 && in_x >= 0.f && in_x <= Wm1) {
        const int y0 = (int)floorf(in_y), y1 = (int)ceilf(in_y), x0 = (int)floorf(in_x), x1 = (int)ceilf(in_x);
        const float ly = __fsub_rn(in_y, (float)y0), lx = __fsub_rn(in_x, (float)x0);
        const int gs = gsrc[g];
        const float tl = masks[((size_t)y0 * MW + x0) * G + gs] ? 1.f : 0.f, tr = masks[((size_t)y0 * MW + x1) * G + gs] ? 1.f : 0.f;
        const float bl = masks[((size_t)y1 * MW + x0) * G + gs] ? 1.f : 0.f, br = masks[((size_t)y1 * MW + x1) * G + gs] ? 1.f : 0.f;
        const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
        const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
        v = rintf(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly)));
      }
    }
    tm[e] = v;
  }
}

size_t det_target_smem() {
  return (size_t)DT_MAX_N * 16 + 2 * (size_t)DT_MAX_N * 8 + 2 * (size_t)DT_MAX_G * 16 + 2 * (size_t)DT_MAX_G * 4 + (size_t)DT_MAX_N * 2;
}

// ------------------------------------------------------------------------------------------------------------------
// PyramidROIAlign backward
// ------------------------------------------------------------------------------------------------------------------
struct RoiBwdParams {
  float* dfeat[4];        // [B,H_l,W_l,C] float32, accumulated into
  int H[4], W[4];
  const float* boxes;     // [B,N,4]
  const int32_t* levels;  // [B,N] (2..5), from the forward
  const __nv_bfloat16* dout;   // [B,N,P,P,C]
  int N, C, P;
};

// One CTA per ROI.  The bilinear weights are separable, so the scatter is done in two steps inside shared memory:
//   T[iy][x][c]  = sum over ix of wx(ix -> x) * dout[iy][ix][c]        (thread = (iy, channel group): owns its row of T)
//   dF[y][x][c] += sum over iy of wy(iy -> y) * T[iy][x][c]            (thread = (y, x, channel group): ONE atomic each)
// over the ROI's footprint (the input pixels its samples touch), 64 channels at a time: an ROI costs fh*fw*C/4 float4
// atomics instead of 4*P*P*C/4 (a 14x14 crop of a 6x6-pixel footprint: 36 instead of 784 per channel group).  ROIs
// whose footprint exceeds 16x16 pixels (long thin boxes on level 2) take the direct path: one atomic per sample corner.
// The sample coordinates repeat the forward (csrc/roialign.cu: in = a1*(D-1) + i*((a2-a1)*(D-1)/(P-1)), outside
// [0, D-1] -> no contribution).
constexpr int RB_FMAX = 16;      // footprint limit of the shared-memory path
constexpr int RB_CH = 64;        // channels per pass

__global__ void __launch_bounds__(256) roialign_backward_kernel(RoiBwdParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_T[];        // [P][RB_FMAX][RB_CH]
  __shared__ int s_lo[2][32], s_hi[2][32];
  __shared__ float s_w[2][32];
  __shared__ int s_min[2], s_max[2];
  const int roi = blockIdx.x, b = roi / p.N;
  const int li = p.levels[roi] - 2;
  const int H = p.H[li], W = p.W[li], C = p.C, P = p.P;
  const float* bp = p.boxes + (size_t)roi * 4;
  if (threadIdx.x < 2) {
    s_min[threadIdx.x] = INT_MAX;
    s_max[threadIdx.x] = -1;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int axis = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i < P) {
      const float a1 = axis == 0 ? bp[0] : bp[1], a2 = axis == 0 ? bp[2] : bp[3];
      const float Dm1 = (float)((axis == 0 ? H : W) - 1);
      const float sc = (P > 1) ? __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(P - 1)) : 0.f;
      const float in = (P > 1) ? __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)i, sc))
                               : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
      int lo = -1, hi = -1;
      float wgt = 0.f;
      if ((in >= 0.f) && (in <= Dm1)) {
        const float f = floorf(in);
        lo = (int)f;
        hi = (int)ceilf(in);
        wgt = __fsub_rn(in, f);
        atomicMin(&s_min[axis], lo);
        atomicMax(&s_max[axis], hi);
      }
      s_lo[axis][i] = lo;
      s_hi[axis][i] = hi;
      s_w[axis][i] = wgt;
    }
  }
  __syncthreads();
  if (s_max[0] < 0 || s_max[1] < 0) return;          // every sample lies outside the map
  const int y0f = s_min[0], x0f = s_min[1];
  const int fh = s_max[0] - y0f + 1, fw = s_max[1] - x0f + 1;
  float* df = p.dfeat[li] + (size_t)b * H * W * C;
  const __nv_bfloat16* g = p.dout + (size_t)roi * P * P * C;
  if (fh > RB_FMAX || fw > RB_FMAX || (C % RB_CH) != 0) {
    // direct path: a thread owns (pixel, 4-channel group) items
    const int cg = C / 4, items = P * P * cg;
    for (int e = threadIdx.x; e < items; e += blockDim.x) {
      const int pix = e / cg, c4 = (e - pix * cg) * 4, iy = pix / P, ix = pix - iy * P;
      const int ya = s_lo[0][iy], yb = s_hi[0][iy], xa = s_lo[1][ix], xb = s_hi[1][ix];
      if ((ya | xa) < 0) continue;
      const float ly = s_w[0][iy], lx = s_w[1][ix];
      const uint2 raw = *reinterpret_cast<const uint2*>(g + (size_t)pix * C + c4);
      const float g0 = __uint_as_float(raw.x << 16), g1 = __uint_as_float(raw.x & 0xffff0000u);
      const float g2 = __uint_as_float(raw.y << 16), g3 = __uint_as_float(raw.y & 0xffff0000u);
      const float wtl = (1.f - lx) * (1.f - ly), wtr = lx * (1.f - ly), wbl = (1.f - lx) * ly, wbr = lx * ly;
      atomicAdd(reinterpret_cast<float4*>(df + ((size_t)ya * W + xa) * C + c4), make_float4(g0 * wtl, g1 * wtl, g2 * wtl, g3 * wtl));
      atomicAdd(reinterpret_cast<float4*>(df + ((size_t)ya * W + xb) * C + c4), make_float4(g0 * wtr, g1 * wtr, g2 * wtr, g3 * wtr));
      atomicAdd(reinterpret_cast<float4*>(df + ((size_t)yb * W + xa) * C + c4), make_float4(g0 * wbl, g1 * wbl, g2 * wbl, g3 * wbl));
      atomicAdd(reinterpret_cast<float4*>(df + ((size_t)yb * W + xb) * C + c4), make_float4(g0 * wbr, g1 * wbr, g2 * wbr, g3 * wbr));
    }
    return;
  }
  constexpr int CG = RB_CH / 4;                     // float4 groups per pass
  float4* T4 = reinterpret_cast<float4*>(s_T);      // [P][RB_FMAX][CG]
  for (int c0 = 0; c0 < C; c0 += RB_CH) {
    // phase 1: row iy of T, one thread per (iy, group)
    for (int e = threadIdx.x; e < P * CG; e += blockDim.x) {
      const int iy = e / CG, cgi = e - iy * CG;
      float4* row = T4 + (size_t)iy * RB_FMAX * CG + cgi;
      for (int x = 0; x < fw; ++x) row[x * CG] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s_lo[0][iy] < 0) continue;
      const __nv_bfloat16* gp = g + (size_t)iy * P * C + c0 + cgi * 4;
      for (int ix = 0; ix < P; ++ix) {
        const int xa = s_lo[1][ix];
        if (xa < 0) continue;
        const int xb = s_hi[1][ix];
        const float lx = s_w[1][ix];
        const uint2 raw = *reinterpret_cast<const uint2*>(gp + (size_t)ix * C);
        const float g0 = __uint_as_float(raw.x << 16), g1 = __uint_as_float(raw.x & 0xffff0000u);
        const float g2 = __uint_as_float(raw.y << 16), g3 = __uint_as_float(raw.y & 0xffff0000u);
        const float wa = 1.f - lx;
        float4 ta = row[(xa - x0f) * CG];
        ta.x += wa * g0; ta.y += wa * g1; ta.z += wa * g2; ta.w += wa * g3;
        row[(xa - x0f) * CG] = ta;
        float4 tb = row[(xb - x0f) * CG];
        tb.x += lx * g0; tb.y += lx * g1; tb.z += lx * g2; tb.w += lx * g3;
        row[(xb - x0f) * CG] = tb;
      }
    }
    __syncthreads();
    // phase 2: one atomic per touched input pixel and channel group
    for (int e = threadIdx.x; e < fh * fw * CG; e += blockDim.x) {
      const int cgi = e % CG, x = (e / CG) % fw, y = e / (CG * fw);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      for (int iy = 0; iy < P; ++iy) {
        const int ya = s_lo[0][iy];
        if (ya < 0) continue;
        const int yb = s_hi[0][iy];
        const float ly = s_w[0][iy];
        float wy = 0.f;
        if (ya - y0f == y) wy += 1.f - ly;
        if (yb - y0f == y) wy += ly;
        if (ya - y0f == y || yb - y0f == y) {
          const float4 t = T4[((size_t)iy * RB_FMAX + x) * CG + cgi];
          acc.x += wy * t.x; acc.y += wy * t.y; acc.z += wy * t.z; acc.w += wy * t.w;
          any = true;
        }
      }
      if (any) atomicAdd(reinterpret_cast<float4*>(df + ((size_t)(y + y0f) * W + (x + x0f)) * C + c0 + cgi * 4), acc);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward "prologue" of a fused conv + BN-affine (+ residual) + ReLU layer: one pass over the incoming gradient
//   dz  = dout * (out > 0)            (gradient before the affine = gradient of the residual input)
//   dzs = dz * scale[c]               (operand of the data- and weight-gradient GEMMs)
//   colsum[c] += sum over rows of dz  (gradient of the shift: bias / beta)
// dout / out / dz / dzs: [rows, C] bf16 (NHWC flattened), C % 8 == 0, C <= 2048.
// ------------------------------------------------------------------------------------------------------------------
struct BwdPrepParams {
  const __nv_bfloat16* dout;
  const __nv_bfloat16* out;      // forward output (post-ReLU) or null: no ReLU
  const float* scale;            // or null
  __nv_bfloat16* dz;             // or null
  __nv_bfloat16* dzs;            // or null
  float* colsum;                 // or null
  long long rows;
  int C;
};

__global__ void __launch_bounds__(256) conv_backward_prep_kernel(BwdPrepParams p) {
  pdl_prologue();
  __shared__ float s_sum[2048];
  const int groups = p.C >> 3;                 // 16-byte channel groups per row
  const int rows_par = 256 / groups;           // rows handled in parallel by the block (>= 1)
  const int cg = threadIdx.x % groups, r0 = threadIdx.x / groups;
  const bool worker = r0 < rows_par;
  for (int c = threadIdx.x; c < p.C; c += 256) s_sum[c] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (worker && p.scale) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = p.scale[cg * 8 + k];
  }
  if (worker) {
    for (long long r = (long long)blockIdx.x * rows_par + r0; r < p.rows; r += (long long)gridDim.x * rows_par) {
      const size_t off = (size_t)r * p.C + (size_t)cg * 8;
      const uint4 g = *reinterpret_cast<const uint4*>(p.dout + off);
      uint4 o = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);     // "positive" when there is no ReLU
      if (p.out) o = *reinterpret_cast<const uint4*>(p.out + off);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
      uint32_t zw[4], sw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float g0 = __uint_as_float(gw[k] << 16), g1 = __uint_as_float(gw[k] & 0xffff0000u);
        const float o0 = __uint_as_float(ow[k] << 16), o1 = __uint_as_float(ow[k] & 0xffff0000u);
        g0 = o0 > 0.f ? g0 : 0.f;
        g1 = o1 > 0.f ? g1 : 0.f;
        acc[2 * k] += g0;
        acc[2 * k + 1] += g1;
        __nv_bfloat162 z = __floats2bfloat162_rn(g0, g1), zs = __floats2bfloat162_rn(g0 * sc[2 * k], g1 * sc[2 * k + 1]);
        zw[k] = *reinterpret_cast<uint32_t*>(&z);
        sw[k] = *reinterpret_cast<uint32_t*>(&zs);
      }
      if (p.dz) *reinterpret_cast<uint4*>(p.dz + off) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
      if (p.dzs) *reinterpret_cast<uint4*>(p.dzs + off) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
    }
    if (p.colsum) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&s_sum[cg * 8 + k], acc[k]);
    }
  }
  __syncthreads();
  if (p.colsum)
    for (int c = threadIdx.x; c < p.C; c += 256) atomicAdd(p.colsum + c, s_sum[c]);
}

// ------------------------------------------------------------------------------------------------------------------
// SGD with momentum, global-norm clipping and the per-tensor L2 regulariser, over one flat buffer
// ------------------------------------------------------------------------------------------------------------------
// segment s covers elements [seg_start[s], seg_start[s+1]) of the flat buffers and carries reg_coef[s] = 2*WEIGHT_DECAY /
// size(w) for kernels / biases, 0 for BatchNorm gamma / beta (mrcnn/model.py:2281-2286: keras l2(wd)(w) / size(w)).
__device__ __forceinline__ int find_segment(const long long* seg_start, int nseg, long long i) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_start[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// pass 1: g += reg_coef * w  (the gradient of the regulariser), sum of squares of the result -> *sumsq (double)
__global__ void __launch_bounds__(256) sgd_norm_kernel(float* __restrict__ grad, const float* __restrict__ w, long long n,
                                                       const long long* __restrict__ seg_start, const float* __restrict__ reg_coef,
                                                       int nseg, float grad_scale, double* sumsq) {
  pdl_prologue();
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    int s = find_segment(seg_start, nseg, i);
    long long send = seg_start[s + 1];
    float coef = reg_coef[s];
    if (i + 4 <= n && i + 4 <= send) {
      float4 g = *reinterpret_cast<float4*>(grad + i);
      const float4 ww = *reinterpret_cast<const float4*>(w + i);
      g.x = g.x * grad_scale + coef * ww.x;
      g.y = g.y * grad_scale + coef * ww.y;
      g.z = g.z * grad_scale + coef * ww.z;
      g.w = g.w * grad_scale + coef * ww.w;
      *reinterpret_cast<float4*>(grad + i) = g;
      acc += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
    } else {
      for (long long k = i; k < i + 4 && k < n; ++k) {
        while (k >= send) {
          ++s;
          send = seg_start[s + 1];
          coef = reg_coef[s];
        }
        const float g = grad[k] * grad_scale + coef * w[k];
        grad[k] = g;
        acc += (double)g * g;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s_acc[8];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_acc[k];
    atomicAdd(sumsq, t);
  }
}

// pass 2: g *= clip / max(norm, clip);  v = momentum*v - lr*g;  w += v;  bf16 operand copy of w refreshed in the same pass
__global__ void __launch_bounds__(256) sgd_apply_kernel(const float* __restrict__ grad, float* __restrict__ w, float* __restrict__ vel,
                                                        __nv_bfloat16* __restrict__ w_bf16, long long n, const double* sumsq,
                                                        float clipnorm, float lr, float momentum) {
  pdl_prologue();
  float scale = 1.f;
  if (clipnorm > 0.f) {
    const float norm = (float)sqrt(*sumsq);
    scale = clipnorm / fmaxf(norm, clipnorm);      // keras clip_norm: g * c / max(n, c)
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i] * scale;
    const float v = momentum * vel[i] - lr * g;
    const float nw = w[i] + v;
    vel[i] = v;
    w[i] = nw;
    if (w_bf16) w_bf16[i] = __float2bfloat16_rn(nw);
  }
}

}  // namespace

extern "C" uint32_t mrcnn_shuffle_key(unsigned long long seed, uint32_t image, uint32_t stream, uint32_t index) {
  return shuffle_key(seed, image, stream, index);
}

extern "C" int mrcnn_detection_targets(const float* proposals, const int32_t* gt_class_ids, const float* gt_boxes,
                                       const uint8_t* gt_masks, int batch, int num_proposals, int max_gt, int mask_src_h,
                                       int mask_src_w, int use_mini_mask, int train_rois, float roi_positive_ratio,
                                       const float* bbox_std_dev, int mask_h, int mask_w, unsigned long long seed,
                                       const unsigned long long* seed_device, float* rois,
                                       int32_t* target_class_ids, float* target_bbox, float* target_mask, int32_t* counts,
                                       void* stream) {
  MRCNN_REQUIRE(proposals && gt_class_ids && gt_boxes && gt_masks && rois && target_class_ids && target_bbox && target_mask &&
                    bbox_std_dev, "detection_targets: null pointer");
  MRCNN_REQUIRE(batch > 0 && num_proposals > 0 && num_proposals <= DT_MAX_N, "detection_targets: proposals per image %d outside [1,%d]",
                num_proposals, DT_MAX_N);
  MRCNN_REQUIRE(max_gt > 0 && max_gt <= DT_MAX_G, "detection_targets: MAX_GT_INSTANCES %d outside [1,%d]", max_gt, DT_MAX_G);
  MRCNN_REQUIRE(train_rois > 0 && mask_h > 0 && mask_w > 0 && mask_src_h > 0 && mask_src_w > 0, "detection_targets: bad shape");
  DetTargetParams p;
  p.proposals = proposals;
  p.gt_class_ids = gt_class_ids;
  p.gt_boxes = gt_boxes;
  p.gt_masks = gt_masks;
  p.N = num_proposals;
  p.G = max_gt;
  p.MH = mask_src_h;
  p.MW = mask_src_w;
  p.T = train_rois;
  p.mask_h = mask_h;
  p.mask_w = mask_w;
  p.use_mini_mask = use_mini_mask;
  p.positive_count = (int)((double)train_rois * (double)roi_positive_ratio);   // int(TRAIN_ROIS_PER_IMAGE * ROI_POSITIVE_RATIO)
  p.inv_ratio = (float)(1.0 / (double)roi_positive_ratio);
  for (int i = 0; i < 4; ++i) p.std_dev[i] = bbox_std_dev[i];
  p.seed = seed;
  p.seed_device = seed_device;
  p.rois = rois;
  p.target_class_ids = target_class_ids;
  p.target_bbox = target_bbox;
  p.target_mask = target_mask;
  p.counts = counts;
  const size_t smem = det_target_smem();
  MRCNN_CHECK_CUDA(cudaFuncSetAttribute(detection_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRCNN_CHECK_CUDA(mrcnn_launch(detection_targets_kernel, dim3(batch), dim3(DT_THREADS), smem, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_pyramid_roi_align_backward(float* const* dfeature_maps, const int* feat_h, const int* feat_w, int channels,
                                                const float* boxes, const int32_t* levels, int batch, int num_boxes,
                                                int pool_size, const void* dpooled_bf16, void* stream) {
  MRCNN_REQUIRE(dfeature_maps && feat_h && feat_w && boxes && levels && dpooled_bf16, "roi_align_backward: null pointer");
  MRCNN_REQUIRE(channels % 4 == 0 && pool_size >= 1 && pool_size <= 32, "roi_align_backward: channels %% 4, pool size <= 32");
  MRCNN_REQUIRE(batch > 0 && num_boxes > 0, "roi_align_backward: empty input");
  RoiBwdParams p;
  for (int i = 0; i < 4; ++i) {
    p.dfeat[i] = dfeature_maps[i];
    p.H[i] = feat_h[i];
    p.W[i] = feat_w[i];
  }
  p.boxes = boxes;
  p.levels = levels;
  p.dout = static_cast<const __nv_bfloat16*>(dpooled_bf16);
  p.N = num_boxes;
  p.C = channels;
  p.P = pool_size;
  const size_t smem = (size_t)pool_size * RB_FMAX * RB_CH * sizeof(float);
  MRCNN_REQUIRE(smem <= 200 * 1024, "roi_align_backward: pool size %d too large", pool_size);
  MRCNN_CHECK_CUDA(cudaFuncSetAttribute(roialign_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRCNN_CHECK_CUDA(mrcnn_launch(roialign_backward_kernel, dim3(batch * num_boxes), dim3(256), smem, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_sgd_step(float* grad, float* weights, float* velocity, void* weights_bf16, long long n,
                              const long long* segment_start, const float* segment_reg_coef, int num_segments,
                              float grad_scale, float clipnorm, float learning_rate, float momentum, double* sumsq_scratch,
                              void* stream) {
  MRCNN_REQUIRE(grad && weights && velocity && segment_start && segment_reg_coef && sumsq_scratch, "sgd_step: null pointer");
  MRCNN_REQUIRE(n > 0 && num_segments > 0, "sgd_step: empty parameter buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MRCNN_CHECK_CUDA(cudaMemsetAsync(sumsq_scratch, 0, sizeof(double), st));
  const int blocks = 148 * 8;
  MRCNN_CHECK_CUDA(mrcnn_launch(sgd_norm_kernel, dim3(blocks), dim3(256), 0, st, grad, (const float*)weights, n, segment_start,
                                segment_reg_coef, num_segments, grad_scale, sumsq_scratch));
  MRCNN_CHECK_CUDA(mrcnn_launch(sgd_apply_kernel, dim3(blocks), dim3(256), 0, st, (const float*)grad, weights, velocity,
                                static_cast<__nv_bfloat16*>(weights_bf16), n, (const double*)sumsq_scratch, clipnorm,
                                learning_rate, momentum));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(2);
  return MRCNN_OK;
}

extern "C" int mrcnn_conv_backward_prep(const void* dout, const void* out, const float* scale, void* dz, void* dzs, float* colsum,
                                        long long rows, int channels, void* stream) {
  MRCNN_REQUIRE(dout && rows > 0, "conv_backward_prep: empty input");
  MRCNN_REQUIRE(channels % 8 == 0 && channels >= 8 && channels <= 2048, "conv_backward_prep: channels %d outside 8..2048 (multiple of 8)",
                channels);
  BwdPrepParams p;
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.out = static_cast<const __nv_bfloat16*>(out);
  p.scale = scale;
  p.dz = static_cast<__nv_bfloat16*>(dz);
  p.dzs = static_cast<__nv_bfloat16*>(dzs);
  p.colsum = colsum;
  p.rows = rows;
  p.C = channels;
  const int rows_par = 256 / (channels / 8);
  long long blocks = (rows + rows_par - 1) / rows_par;
  if (blocks > 148 * 8) blocks = 148 * 8;
  MRCNN_CHECK_CUDA(mrcnn_launch(conv_backward_prep_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}
