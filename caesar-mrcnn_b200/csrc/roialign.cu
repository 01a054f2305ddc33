// PyramidROIAlign — level-routed tf.image.crop_and_resize(bilinear), one CTA per ROI, NHWC features,
// 16-byte vector loads (8 bf16 / 4 f32 channels per lane), streaming 16-byte stores.
// Replaces mrcnn/model.py:428-534 (PyramidROIAlign.call: 4x tf.where + gather_nd + crop_and_resize,
// concat, top_k re-sort) and :413-423 (log2_graph).  The reference's "route, crop, restore order"
// is the identity out[b,n] = crop(P_level(b,n)[b], boxes[b,n]), which is what is computed here.
// Contract: ROI levels bit-exact; float32 features -> bit-exact samples (no FMA contraction);
// bf16 features -> float32 lerp of the bf16 values, one rounding to bf16 on store.
#include <stdlib.h>
#include "box_ops.cuh"
#include "mrcnn_b200.h"

void mrcnn_count_launch(unsigned long long n);

namespace {

struct RoiParams {
  const void* feat[4];  // P2..P5, [B,H_l,W_l,C]
  int H[4], W[4];
  const float* boxes;   // [B,N,box_stride] normalised (y1,x1,y2,x2,...)
  int box_stride;
  int N, C, P;
  float image_area;
  void* out;            // [B,N,P,P,C]
  int32_t* levels;      // [B,N] or null
  int levels_ready;     // levels[] was filled by roi_levels_kernel (all ROIs in parallel) before this launch
  unsigned long long one2;  // (1.0f, 1.0f) as a packed fp32 pair, see f2_add
  int valid_col;        // >= 0: box column holding the class id; a ROI whose class id is 0 (zero padding of the detections)
                        // is skipped and its output rows are left untouched.  -1: every ROI is pooled
};

// mrcnn/model.py:465-477 — level = min(5, max(2, 4 + int32(round(log2(sqrt(h*w)/(224/sqrt(area)))))))
// log convention: double log, one rounding to float (oracle/graph_layers.py: log_f32).
__device__ __forceinline__ int roi_level(float y1, float x1, float y2, float x2, float image_area) {
  const float h = __fsub_rn(y2, y1);
  const float w = __fsub_rn(x2, x1);
  const float denom = __fdiv_rn(224.0f, __fsqrt_rn(image_area));
  const float v = __fdiv_rn(__fsqrt_rn(__fmul_rn(h, w)), denom);
  const float lg = (float)log((double)v);
  const float ln2 = (float)log(2.0);
  const float r = rintf(__fdiv_rn(lg, ln2));  // tf.round: half to even
  // tf.cast(float -> int32) on x86 (cvttss2si): NaN / out-of-range -> INT_MIN
  int ri;
  if (!(fabsf(r) < 2147483648.0f)) ri = INT_MIN; else ri = (int)r;
  const int lvl = (int)((unsigned)4 + (unsigned)ri);  // int32 wrap-around
  return min(5, max(2, lvl));
}

template <typename T> struct Vec;  // 16-byte channel vector, kept raw in registers until the lerp
template <> struct Vec<float> {
  static constexpr int N = 4;
  typedef float4 Raw;
  __device__ static Raw load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static void unpack(const Raw& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ static void store(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  typedef uint4 Raw;
  __device__ static Raw load(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void unpack(const Raw& t, float* v) {
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3)));
  }
};

// Packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): every lane of a pair is an IEEE round-to-nearest fp32
// operation, so the samples are bit-identical to the scalar left-to-right evaluation.  ptxas contracts a packed
// mul.rn + add.rn pair into one FFMA2 even under -fmad=false (it does not for scalar fp32), which would round
// once instead of twice; the addition is therefore issued as fma(t, one, a) with `one` = (1.0f, 1.0f) taken
// from the kernel parameters (opaque to the optimizer): t*1 is exact, so the result is exactly t + a.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b, uint64_t one) {   // a + b
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(b), "l"(one), "l"(a));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t b, uint64_t a) {   // b - a
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
  return d;
}
__device__ __forceinline__ uint64_t f2_from_bf16x2(uint32_t w) {
  return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(uint64_t v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// tf.image.crop_and_resize bilinear sample of two channels: top = tl + (tr - tl)*lx; bot likewise; top + (bot - top)*ly
__device__ __forceinline__ uint32_t lerp_bf16x2(uint32_t tl, uint32_t tr, uint32_t bl, uint32_t br, uint64_t lx, uint64_t ly,
                                                uint64_t one) {
  const uint64_t a = f2_from_bf16x2(tl), b = f2_from_bf16x2(tr), c = f2_from_bf16x2(bl), d = f2_from_bf16x2(br);
  const uint64_t top = f2_add(a, f2_mul(f2_sub(b, a), lx), one);
  const uint64_t bot = f2_add(c, f2_mul(f2_sub(d, c), lx), one);
  return f2_to_bf16x2(f2_add(top, f2_mul(f2_sub(bot, top), ly), one));
}

// One CTA per ROI.  Thread 0 computes the pyramid level; the P sample rows / columns of the crop are
// then computed once per ROI into shared memory (tf.image.crop_and_resize coordinates, float32,
// left-to-right evaluation) as element offsets + lerp weights; each warp takes output pixels
// round-robin with one lane per 16-byte channel vector, so a pixel is 4 coalesced 512-byte reads
// (bf16, C=256) and one coalesced 512-byte streaming store, with no per-pixel address arithmetic
// beyond four integer adds.
constexpr int ROI_MAX_P = 32;

// 7 warps: the 49 (7x7) and 196 (14x14) output pixels of the two heads split evenly over them
constexpr int ROI_THREADS = 224;

template <typename T>
__global__ void __launch_bounds__(ROI_THREADS) roialign_kernel(RoiParams p) {
  pdl_prologue();
  constexpr int VN = Vec<T>::N;
  __shared__ int s_lo[2][ROI_MAX_P], s_hi[2][ROI_MAX_P];   // element offsets: [0] rows (t*W*C, b*W*C), [1] cols (l*C, r*C); -1 = outside
  __shared__ float s_w[2][ROI_MAX_P];                      // lerp weights
  __shared__ int s_level;
  const int roi = blockIdx.x;  // b*N + n
  const int b = roi / p.N;
  const float* bp = p.boxes + (size_t)roi * p.box_stride;
  if (p.valid_col >= 0 && bp[p.valid_col] == 0.f) return;
  const float y1 = bp[0], x1 = bp[1], y2 = bp[2], x2 = bp[3];
  int li;
  if (p.levels_ready) {
    li = p.levels[roi] - 2;
  } else {
    // one thread evaluates the double-precision log; the whole CTA waits (only used when no level buffer is given)
    if (threadIdx.x == 0) {
      const int lv = roi_level(y1, x1, y2, x2, p.image_area);
      s_level = lv;
      if (p.levels) p.levels[roi] = lv;
    }
    __syncthreads();
    li = s_level - 2;
  }
  const int H = p.H[li], W = p.W[li], C = p.C, P = p.P;
  const T* feat = static_cast<const T*>(p.feat[li]) + (size_t)b * H * W * C;
  T* out = static_cast<T*>(p.out) + (size_t)roi * P * P * C;
  if (threadIdx.x < 2 * ROI_MAX_P) {
    const int axis = threadIdx.x / ROI_MAX_P, i = threadIdx.x % ROI_MAX_P;
    if (i < P) {
      const float a1 = axis == 0 ? y1 : x1, a2 = axis == 0 ? y2 : x2;
      const float Dm1 = (float)((axis == 0 ? H : W) - 1);
      const float sc = (P > 1) ? __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(P - 1)) : 0.f;
      const float in = (P > 1) ? __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)i, sc))
                               : __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
      int lo = -1, hi = -1;
      float wgt = 0.f;
      if ((in >= 0.f) && (in <= Dm1)) {
        const float f = floorf(in);
        const int mul = axis == 0 ? W * C : C;
        lo = (int)f * mul;
        hi = (int)ceilf(in) * mul;
        wgt = __fsub_rn(in, f);
      }
      s_lo[axis][i] = lo;
      s_hi[axis][i] = hi;
      s_w[axis][i] = wgt;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if constexpr (sizeof(T) == 2) {
    if (C == 256) {
      // bf16, 256 channels (the engine's pyramid): one lane per 16-byte vector covers a pixel with one warp pass;
      // two output pixels per iteration keep 8 independent 16-byte gathers in flight per lane
      // 32-bit byte offsets from one 64-bit lane base, (iy, ix) advanced incrementally: no integer division and no
      // 64-bit index arithmetic per gather
      const char* fb = reinterpret_cast<const char*>(feat) + lane * 16;
      char* ob = reinterpret_cast<char*>(out) + lane * 16;
      const int npix = P * P;
      const uint64_t one2 = p.one2;
      struct Px {
        uint4 tl, tr, bl, br;
        uint64_t lx2, ly2;
        bool valid;
      };
      auto sample = [&](int iy, int ix, Px& q) {
        const int ro0 = s_lo[0][iy], ro1 = s_hi[0][iy], co0 = s_lo[1][ix], co1 = s_hi[1][ix];
        const float ly = s_w[0][iy], lx = s_w[1][ix];
        q.lx2 = f2_pack(lx, lx);
        q.ly2 = f2_pack(ly, ly);
        q.valid = (ro0 | co0) >= 0;
        if (q.valid) {
          q.tl = __ldg(reinterpret_cast<const uint4*>(fb + 2u * (uint32_t)(ro0 + co0)));
          q.tr = __ldg(reinterpret_cast<const uint4*>(fb + 2u * (uint32_t)(ro0 + co1)));
          q.bl = __ldg(reinterpret_cast<const uint4*>(fb + 2u * (uint32_t)(ro1 + co0)));
          q.br = __ldg(reinterpret_cast<const uint4*>(fb + 2u * (uint32_t)(ro1 + co1)));
        }
      };
      auto finish = [&](int pix, const Px& q) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (q.valid) {
          o.x = lerp_bf16x2(q.tl.x, q.tr.x, q.bl.x, q.br.x, q.lx2, q.ly2, one2);
          o.y = lerp_bf16x2(q.tl.y, q.tr.y, q.bl.y, q.br.y, q.lx2, q.ly2, one2);
          o.z = lerp_bf16x2(q.tl.z, q.tr.z, q.bl.z, q.br.z, q.lx2, q.ly2, one2);
          o.w = lerp_bf16x2(q.tl.w, q.tr.w, q.bl.w, q.br.w, q.lx2, q.ly2, one2);
        }
        __stcs(reinterpret_cast<uint4*>(ob + 512u * (uint32_t)pix), o);
      };
      // pixel of this warp and the step to its next one (nwarps pixels further), as (row, column) pairs
      int iy = warp / P, ix = warp - iy * P;
      const int dy = nwarps / P, dx = nwarps - dy * P;
      auto advance = [&](int& y, int& x) {
        y += dy;
        x += dx;
        if (x >= P) {
          x -= P;
          ++y;
        }
      };
      int pix = warp;
      for (; pix + nwarps < npix; pix += 2 * nwarps) {
        Px q0, q1;
        int iy1 = iy, ix1 = ix;
        advance(iy1, ix1);
        sample(iy, ix, q0);
        sample(iy1, ix1, q1);
        finish(pix, q0);
        finish(pix + nwarps, q1);
        iy = iy1;
        ix = ix1;
        advance(iy, ix);
      }
      if (pix < npix) {
        Px q0;
        sample(iy, ix, q0);
        finish(pix, q0);
      }
      return;
    }
  }
  int iy = warp / P, ix = warp - iy * P;
  for (int pix = warp; pix < P * P; pix += nwarps) {
    const int ro0 = s_lo[0][iy], ro1 = s_hi[0][iy], co0 = s_lo[1][ix], co1 = s_hi[1][ix];
    const float ly = s_w[0][iy], lx = s_w[1][ix];
    const bool valid = (ro0 >= 0) && (co0 >= 0);
    T* po = out + pix * C;
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float o[VN];
      if (valid) {
        const T* fb = feat + c0;
        const typename Vec<T>::Raw rtl = Vec<T>::load(fb + ro0 + co0);
        const typename Vec<T>::Raw rtr = Vec<T>::load(fb + ro0 + co1);
        const typename Vec<T>::Raw rbl = Vec<T>::load(fb + ro1 + co0);
        const typename Vec<T>::Raw rbr = Vec<T>::load(fb + ro1 + co1);
        float a[VN], bq[VN], c[VN], d[VN];
        Vec<T>::unpack(rtl, a);
        Vec<T>::unpack(rtr, bq);
        Vec<T>::unpack(rbl, c);
        Vec<T>::unpack(rbr, d);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          const float top = __fadd_rn(a[k], __fmul_rn(__fsub_rn(bq[k], a[k]), lx));
          const float bot = __fadd_rn(c[k], __fmul_rn(__fsub_rn(d[k], c[k]), lx));
          o[k] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
        }
      } else {
#pragma unroll
        for (int k = 0; k < VN; ++k) o[k] = 0.f;
      }
      Vec<T>::store(po + c0, o);
    }
    ix += nwarps;
    while (ix >= P) {
      ix -= P;
      ++iy;
    }
  }
}

// ---- row-per-warp variant for the engine's case (bf16, C = 256, P = 7 | 14) ----------------------------------------
// The generic kernel above spends ~146 SASS instructions per output pixel and lane, of which only 72 are the
// bilinear arithmetic (32 bf16->f32 unpacks, 36 packed fp32 ops that must not be contracted, 4 packs): the rest is
// per-pixel table lookups, 64-bit address assembly and loop bookkeeping, and the kernel is issue-bound (ncu: 70-77 %
// issue-slot utilisation, DRAM traffic = the algorithmic bytes).  Here one warp owns a whole output ROW of the crop:
// the two source-row pointers and the y weight are set up once per row, the column loop is fully unrolled (P is a
// template parameter) with one 16-byte table entry per column, every gather address is one 64-bit add of a
// precomputed BYTE offset, the store offset is an immediate, and the loads of pixel ix+1 are issued before the
// arithmetic of pixel ix.  Same operations in the same order on the same values: bit-identical output.
struct __align__(16) AxisEnt {
  uint32_t lo, hi;      // byte offsets of the floor / ceil source row (x W*C*2) or column (x C*2); valid only if ok
  float w;              // lerp weight
  int ok;               // sample lies inside [0, D-1] (tf.image.crop_and_resize extrapolates to 0 outside)
};

template <int P, int MINB, bool PREFETCH = true>
__global__ void __launch_bounds__(ROI_THREADS, MINB) roialign_rows_kernel(RoiParams p) {
  pdl_prologue();
  __shared__ AxisEnt s_ax[2][P];      // [0] rows, [1] columns
  const int roi = blockIdx.x;         // b*N + n
  const int b = roi / p.N;
  const float* bp = p.boxes + (size_t)roi * p.box_stride;
  if (p.valid_col >= 0 && bp[p.valid_col] == 0.f) return;
  const int li = p.levels[roi] - 2;
  const int H = p.H[li], W = p.W[li];
  constexpr int C = 256;
  if (threadIdx.x < 2 * P) {
    const int axis = threadIdx.x / P, i = threadIdx.x - axis * P;
    const float a1 = axis == 0 ? bp[0] : bp[1], a2 = axis == 0 ? bp[2] : bp[3];
    const float Dm1 = (float)((axis == 0 ? H : W) - 1);
    const float sc = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(P - 1));
    const float in = __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)i, sc));
    AxisEnt e;
    e.lo = 0u; e.hi = 0u; e.w = 0.f; e.ok = 0;
    if ((in >= 0.f) && (in <= Dm1)) {
      const float f = floorf(in);
      const uint32_t mul = (axis == 0 ? (uint32_t)W * C : (uint32_t)C) * 2u;
      e.lo = (uint32_t)(int)f * mul;
      e.hi = (uint32_t)(int)ceilf(in) * mul;
      e.w = __fsub_rn(in, f);
      e.ok = 1;
    }
    s_ax[axis][i] = e;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const char* fb = reinterpret_cast<const char*>(p.feat[li]) + ((size_t)b * H * W * C) * 2 + lane * 16;
  char* ob = reinterpret_cast<char*>(p.out) + ((size_t)roi * P * P * C) * 2 + lane * 16;
  const uint64_t one2 = p.one2;
  for (int iy = warp; iy < P; iy += ROI_THREADS / 32) {
    const AxisEnt re = s_ax[0][iy];
    char* orow = ob + (size_t)iy * (P * C * 2);
    if (!re.ok) {
#pragma unroll
      for (int ix = 0; ix < P; ++ix) __stcs(reinterpret_cast<uint4*>(orow + ix * (C * 2)), make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    const char* r0 = fb + re.lo;
    const char* r1 = fb + re.hi;
    const uint64_t ly2 = f2_pack(re.w, re.w);
    // gathers are unconditional (a column outside the map has offsets 0: a valid address whose data is ignored), so
    // the loads of the next pixel need no predicate and the compiler can rename registers instead of copying them
    if constexpr (!PREFETCH) {
      // high-occupancy form: no software double buffer (16 registers less); the other resident warps hide the gathers
#pragma unroll
      for (int ix = 0; ix < P; ++ix) {
        const AxisEnt e = s_ax[1][ix];
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (e.ok) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(r0 + e.lo));
          const uint4 b2 = __ldg(reinterpret_cast<const uint4*>(r0 + e.hi));
          const uint4 c = __ldg(reinterpret_cast<const uint4*>(r1 + e.lo));
          const uint4 d = __ldg(reinterpret_cast<const uint4*>(r1 + e.hi));
          const uint64_t lx2 = f2_pack(e.w, e.w);
          o.x = lerp_bf16x2(a.x, b2.x, c.x, d.x, lx2, ly2, one2);
          o.y = lerp_bf16x2(a.y, b2.y, c.y, d.y, lx2, ly2, one2);
          o.z = lerp_bf16x2(a.z, b2.z, c.z, d.z, lx2, ly2, one2);
          o.w = lerp_bf16x2(a.w, b2.w, c.w, d.w, lx2, ly2, one2);
        }
        __stcs(reinterpret_cast<uint4*>(orow + ix * (C * 2)), o);
      }
      continue;
    }
    AxisEnt ce = s_ax[1][0];
    uint4 tl = __ldg(reinterpret_cast<const uint4*>(r0 + ce.lo));
    uint4 tr = __ldg(reinterpret_cast<const uint4*>(r0 + ce.hi));
    uint4 bl = __ldg(reinterpret_cast<const uint4*>(r1 + ce.lo));
    uint4 br = __ldg(reinterpret_cast<const uint4*>(r1 + ce.hi));
#pragma unroll
    for (int ix = 0; ix < P; ++ix) {
      uint4 ntl, ntr, nbl, nbr;
      AxisEnt ne;
      if (ix + 1 < P) {
        ne = s_ax[1][ix + 1];
        ntl = __ldg(reinterpret_cast<const uint4*>(r0 + ne.lo));
        ntr = __ldg(reinterpret_cast<const uint4*>(r0 + ne.hi));
        nbl = __ldg(reinterpret_cast<const uint4*>(r1 + ne.lo));
        nbr = __ldg(reinterpret_cast<const uint4*>(r1 + ne.hi));
      }
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (ce.ok) {
        const uint64_t lx2 = f2_pack(ce.w, ce.w);
        o.x = lerp_bf16x2(tl.x, tr.x, bl.x, br.x, lx2, ly2, one2);
        o.y = lerp_bf16x2(tl.y, tr.y, bl.y, br.y, lx2, ly2, one2);
        o.z = lerp_bf16x2(tl.z, tr.z, bl.z, br.z, lx2, ly2, one2);
        o.w = lerp_bf16x2(tl.w, tr.w, bl.w, br.w, lx2, ly2, one2);
      }
      __stcs(reinterpret_cast<uint4*>(orow + ix * (C * 2)), o);
      if (ix + 1 < P) {
        ce = ne;
        tl = ntl; tr = ntr; bl = nbl; br = nbr;
      }
    }
  }
}

__global__ void roi_levels_kernel(const float* boxes, int box_stride, int n, float image_area, int32_t* levels) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float* bx = boxes + (size_t)i * box_stride;
    levels[i] = roi_level(bx[0], bx[1], bx[2], bx[3], image_area);
  }
}

}  // namespace

extern "C" int mrcnn_roi_levels(const float* boxes, int num_boxes, float image_area, int32_t* levels, void* stream) {
  MRCNN_REQUIRE(boxes && levels && num_boxes > 0, "roi_levels: bad arguments");
  MRCNN_CHECK_CUDA(mrcnn_launch(roi_levels_kernel, dim3(ceil_div(num_boxes, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), boxes, 4, num_boxes, image_area, levels));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  return MRCNN_OK;
}

int launch_pyramid_roi_align(const void* const* feature_maps, const int* feat_h, const int* feat_w, int channels,
                             int dtype, const float* boxes, int box_stride, int batch, int num_boxes, int pool_size,
                             float image_area, void* pooled, int32_t* levels, cudaStream_t st, int valid_col) {
  MRCNN_REQUIRE(feature_maps && feat_h && feat_w && boxes && pooled, "pyramid_roi_align: null pointer");
  MRCNN_REQUIRE(valid_col < box_stride, "pyramid_roi_align: valid_col outside the box rows");
  MRCNN_REQUIRE(batch > 0 && num_boxes > 0 && pool_size >= 1, "pyramid_roi_align: empty input");
  MRCNN_REQUIRE(pool_size <= ROI_MAX_P, "pyramid_roi_align: pool_size %d > %d", pool_size, ROI_MAX_P);
  MRCNN_REQUIRE(dtype == MRCNN_DTYPE_F32 || dtype == MRCNN_DTYPE_BF16, "pyramid_roi_align: dtype must be f32 or bf16");
  const int vn = dtype == MRCNN_DTYPE_F32 ? 4 : 8;
  MRCNN_REQUIRE(channels % vn == 0, "pyramid_roi_align: channels=%d must be a multiple of %d", channels, vn);
  RoiParams p;
  for (int i = 0; i < 4; ++i) {
    MRCNN_REQUIRE(feature_maps[i] && feat_h[i] >= 1 && feat_w[i] >= 1, "pyramid_roi_align: bad level %d", i + 2);
    p.feat[i] = feature_maps[i];
    p.H[i] = feat_h[i];
    p.W[i] = feat_w[i];
  }
  p.boxes = boxes;
  p.box_stride = box_stride;
  p.N = num_boxes;
  p.C = channels;
  p.P = pool_size;
  p.image_area = image_area;
  p.out = pooled;
  p.levels = levels;
  p.levels_ready = 0;
  p.valid_col = valid_col;
  if (levels) {   // all levels in one parallel pass instead of one serial double-log per ROI CTA
    MRCNN_CHECK_CUDA(mrcnn_launch(roi_levels_kernel, dim3(ceil_div(batch * num_boxes, 256)), dim3(256), 0, st, boxes, box_stride, batch * num_boxes, image_area, levels));
    MRCNN_CHECK_CUDA(cudaGetLastError());
    mrcnn_count_launch(1);
    p.levels_ready = 1;
  }
  p.one2 = 0x3f8000003f800000ull;
  static const bool rows_variant = !(getenv("MRCNN_B200_ROIALIGN_ROWS") && getenv("MRCNN_B200_ROIALIGN_ROWS")[0] == '0');
  if (dtype == MRCNN_DTYPE_BF16 && channels == 256 && p.levels_ready && rows_variant && (pool_size == 7 || pool_size == 14)) {
    // occupancy is what this issue-bound kernel lacked (ncu: 62 % issue-slot use at 4 CTAs per SM): 7 CTAs per SM without the
    // software double buffer for the 7x7 pool (0.55 -> 0.47 ms per 64 000 ROIs), 6 with it for 14x14 (0.197 -> 0.188 ms);
    // 8 per SM and the double-buffered form at 7 per SM spill and lose
    if (pool_size == 7)
      MRCNN_CHECK_CUDA(mrcnn_launch(roialign_rows_kernel<7, 7, false>, dim3(batch * num_boxes), dim3(ROI_THREADS), 0, st, p));
    else
      MRCNN_CHECK_CUDA(mrcnn_launch(roialign_rows_kernel<14, 6, true>, dim3(batch * num_boxes), dim3(ROI_THREADS), 0, st, p));
  } else if (dtype == MRCNN_DTYPE_F32)
    MRCNN_CHECK_CUDA(mrcnn_launch(roialign_kernel<float>, dim3(batch * num_boxes), dim3(ROI_THREADS), 0, st, p));
  else
    MRCNN_CHECK_CUDA(mrcnn_launch(roialign_kernel<__nv_bfloat16>, dim3(batch * num_boxes), dim3(ROI_THREADS), 0, st, p));
  MRCNN_CHECK_CUDA(cudaGetLastError());
  mrcnn_count_launch(1);
  return MRCNN_OK;
}

extern "C" int mrcnn_pyramid_roi_align(const void* const* feature_maps, const int* feat_h, const int* feat_w,
                                       int channels, int dtype, const float* boxes, int batch, int num_boxes,
                                       int pool_size, float image_area, void* pooled, int32_t* levels,
                                       void* stream) {
  return launch_pyramid_roi_align(feature_maps, feat_h, feat_w, channels, dtype, boxes, 4, batch, num_boxes, pool_size,
                                  image_area, pooled, levels, static_cast<cudaStream_t>(stream), -1);
}
